set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
P="timeout 300 python tools/probe.py"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/f1_pytest.log; cat $O/f1_pytest.log
$P 1 3 PROBE_CONST=1 > $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_FORCE_RXY=4 >> $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_FORCE_RXY=7 >> $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_FORCE_RXY=7 SWEEPTT_INNER=1 >> $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_FORCE_RXY=4 SWEEPTT_INNER=2 >> $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_INNER=2 >> $O/f1_probe.log 2>&1
$P 1 3 PROBE_CONST=1 SWEEPTT_INNER=8 >> $O/f1_probe.log 2>&1
$P 4 3 >> $O/f1_probe.log 2>&1
$P 4 3 SWEEPTT_FORCE_RXY=7 >> $O/f1_probe.log 2>&1
cat $O/f1_probe.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 > $O/f1_bench.json 2> $O/f1_bench.err; tail -c 300 $O/f1_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:relax_tiled --launch-skip 2 -c 1 -f -o $O/r02_relax_c2_final python tools/probe.py 4 > $O/f1_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/f1_ncu_bench.log 2>&1
timeout 600 python tools/cli_e2e.py 2>&1 | tail -3
