set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:relax_tiled --launch-skip 2 -c 1 -f -o $O/r02_relax_5fs python tools/probe.py 4 5 > $O/f8_ncu5.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:relax_tiled --launch-skip 140 -c 1 -f -o $O/r02_relax_3fs python tools/probe.py 4 3 SWEEPTT_LOOP=batched > $O/f8_ncu3.log 2>&1
tail -2 $O/f8_ncu5.log $O/f8_ncu3.log
