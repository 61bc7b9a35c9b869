set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 2 > $O/m2f.log 2>&1
$G 2 SWEEPTT_FRONT_SLACK=0 >> $O/m2f.log 2>&1
$G 2 SWEEPTT_FRONT_SLACK=2 >> $O/m2f.log 2>&1
$G 2 SWEEPTT_FRONT_SLACK=100 >> $O/m2f.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=4 >> $O/m2f.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=2 >> $O/m2f.log 2>&1
$G 2 8 >> $O/m2f.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -3 >> $O/m2f.log
grep -E "^\[|sweeptt\]|passed|failed|cases" $O/m2f.log
