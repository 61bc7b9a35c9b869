set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; free -g | head -2
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/c1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err
P="timeout 300 python tools/probe.py"
$P 111 >> $O/c1_probe.log 2>&1
$P 111 SWEEPTT_WAVE=0 >> $O/c1_probe.log 2>&1
$P 111 SWEEPTT_WAVE=4 >> $O/c1_probe.log 2>&1
$P 111 SWEEPTT_WAVE_STREAMS=2 >> $O/c1_probe.log 2>&1
$P 111 SWEEPTT_WAVE_STREAMS=4 >> $O/c1_probe.log 2>&1
$P 111 SWEEPTT_NO_XCLIP=1 >> $O/c1_probe.log 2>&1
$P 14 >> $O/c1_probe.log 2>&1
$P 14 SWEEPTT_WAVE_STREAMS=2 >> $O/c1_probe.log 2>&1
$P 8 >> $O/c1_probe.log 2>&1
$P 4 >> $O/c1_probe.log 2>&1
$P 4 SWEEPTT_NO_XCLIP=1 >> $O/c1_probe.log 2>&1
$P 1 >> $O/c1_probe.log 2>&1
for b in 2 4 8 16; do $P 1 3 PROBE_CONST=1 SWEEPTT_BUCKET=$b >> $O/c1_probe.log 2>&1; done
$P 1 3 PROBE_CONST=1 SWEEPTT_PERSIST=1 >> $O/c1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_PERSIST=1 SWEEPTT_BUCKET=8 >> $O/c1_probe.log 2>&1
$P 4 5 >> $O/c1_probe.log 2>&1
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_tx4.so $P 4 TX=4 >> $O/c1_probe.log 2>&1
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_tx4.so $P 8 TX=4 >> $O/c1_probe.log 2>&1
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_dbg.so timeout 900 python tools/random_parity.py 7000 45 > $O/c1_dbg_bounds.log 2>&1
timeout 600 python tools/cli_e2e.py > $O/c1_cli.log 2>&1
timeout 1200 python tools/legacy_gpu.py > $O/c1_legacy.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/c1_ref.json 2> $O/c1_ref.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:relax_tiled --launch-skip 2 -c 1 -f -o $O/r02_relax_c2 python tools/probe.py 4 > $O/c1_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/c1_ncu_bench.log 2>&1
cat $O/c1_pytest.log $O/c1_probe.log; tail -3 $O/c1_dbg_bounds.log; tail -2 $O/c1_cli.log; tail -3 $O/c1_legacy.log; tail -c 600 $O/c1_bench.err
