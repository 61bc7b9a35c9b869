set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
L=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 1 > $O/m2h.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_noearly.so $G 1 NOEARLY=1 >> $O/m2h.log 2>&1
$G 2 >> $O/m2h.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_noearly.so $G 2 NOEARLY=1 >> $O/m2h.log 2>&1
$G 2 SWEEPTT_BLOCK_TILES=4 >> $O/m2h.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_noearly.so $G 2 NOEARLY=1 SWEEPTT_BLOCK_TILES=4 >> $O/m2h.log 2>&1
timeout 300 python tools/probe.py 1 3 PROBE_CONST=1 >> $O/m2h.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_noearly.so timeout 300 python tools/probe.py 1 3 PROBE_CONST=1 NOEARLY=1 >> $O/m2h.log 2>&1
timeout 300 python tools/probe.py 111 SWEEPTT_WAVE=0 >> $O/m2h.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_noearly.so timeout 300 python tools/probe.py 111 SWEEPTT_WAVE=0 NOEARLY=1 >> $O/m2h.log 2>&1
grep -E "^\[" $O/m2h.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
