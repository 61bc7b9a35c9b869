set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/e2e_probe.py 14 2> $O/c3_e2e14.log
python tools/e2e_probe.py 111 2> $O/c3_e2e111.log
grep -E "rep 3|timing" $O/c3_e2e14.log | tail -8; grep -E "rep" $O/c3_e2e111.log
P="timeout 300 python tools/probe.py"
$P 111 > $O/c3_probe.log 2>&1
$P 111 SWEEPTT_WAVE_STREAMS=1 >> $O/c3_probe.log 2>&1
$P 14 >> $O/c3_probe.log 2>&1
$P 4 >> $O/c3_probe.log 2>&1
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_nopin.so $P 4 NOPIN=1 >> $O/c3_probe.log 2>&1
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_nopin.so $P 111 NOPIN=1 SWEEPTT_WAVE_STREAMS=1 >> $O/c3_probe.log 2>&1
cat $O/c3_probe.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/c3_bench.json 2> $O/c3_bench.err; tail -c 300 $O/c3_bench.err
timeout 600 python tools/cli_e2e.py 2>&1 | tail -3
