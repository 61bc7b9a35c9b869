set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 1 > $O/m2d.log 2>&1
$G 2 >> $O/m2d.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=8 >> $O/m2d.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=16 >> $O/m2d.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=2 >> $O/m2d.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=8 SWEEPTT_BIAS=8,2,14,8,2,14 >> $O/m2d.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=8 SWEEPTT_NO_GLOBAL_KMIN=1 >> $O/m2d.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -3 >> $O/m2d.log
grep -E "^\[|sweeptt\]|passed|failed" $O/m2d.log
