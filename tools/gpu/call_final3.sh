set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:relax_tiled --launch-skip 2 -c 1 -f -o $O/r02_relax_c3_wave8 python tools/probe.py 8 > $O/f3_ncu.log 2>&1
tail -3 $O/f3_ncu.log
