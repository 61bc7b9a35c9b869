set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
P="timeout 300 python tools/probe.py"
L=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp
$P 4 > $O/c4_probe.log 2>&1
$P 8 >> $O/c4_probe.log 2>&1
$P 111 >> $O/c4_probe.log 2>&1
for nw in 20 24; do
  SWEEPTT_LIB=$L/libsweeptt_nw$nw.so $P 4 NW=$nw >> $O/c4_probe.log 2>&1
  SWEEPTT_LIB=$L/libsweeptt_nw$nw.so $P 8 NW=$nw >> $O/c4_probe.log 2>&1
  SWEEPTT_LIB=$L/libsweeptt_nw$nw.so $P 111 NW=$nw >> $O/c4_probe.log 2>&1
  SWEEPTT_LIB=$L/libsweeptt_nw$nw.so $P 1 NW=$nw >> $O/c4_probe.log 2>&1
  SWEEPTT_LIB=$L/libsweeptt_nw$nw.so $P 4 5 NW=$nw >> $O/c4_probe.log 2>&1
done
SWEEPTT_LIB=$L/libsweeptt_nw20.so $P 8 NW=20 SWEEPTT_BIAS=8,2,14,10,2,40 >> $O/c4_probe.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_nw20.so $P 8 NW=20 SWEEPTT_BIAS=8,2,14,10,2,90 >> $O/c4_probe.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_nw20.so $P 8 NW=20 SWEEPTT_BIAS=8,2,14,16,4,60 >> $O/c4_probe.log 2>&1
SWEEPTT_LIB=$L/libsweeptt_nw20.so $P 8 NW=20 SWEEPTT_COLCOST=2.5 >> $O/c4_probe.log 2>&1
cat $O/c4_probe.log | cut -c1-260
SWEEPTT_LIB=$L/libsweeptt_nw20.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
