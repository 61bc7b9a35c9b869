set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 1 > $O/m2b.log 2>&1
$G 2 >> $O/m2b.log 2>&1
$G 2 SWEEPTT_NO_GLOBAL_KMIN=1 >> $O/m2b.log 2>&1
for t in 1 2 8 16; do $G 2 SWEEPTT_BLOCK_TILES=$t >> $O/m2b.log 2>&1; done
for k in 1 2 8 16; do $G 2 $k >> $O/m2b.log 2>&1; done
$G 2 SWEEPTT_BUCKET=4 >> $O/m2b.log 2>&1
$G 2 SWEEPTT_BUCKET=1 >> $O/m2b.log 2>&1
grep -E "^\[|sweeptt\]" $O/m2b.log
