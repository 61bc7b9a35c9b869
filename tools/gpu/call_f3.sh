set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
P="timeout 300 python tools/probe.py"
rm -f $O/f3_probe.log
for b in "8,2,14,10,2,60" "8,2,14,10,2,40" "8,2,14,10,2,80" "8,2,14,14,2,60" "8,2,14,6,2,60" "8,2,14,10,6,60" "8,2,14,10,0,60"; do $P 8 SWEEPTT_BIAS=$b >> $O/f3_probe.log 2>&1; done
for l in 4 8 16 32; do $P 8 SWEEPTT_LOOKAHEAD=$l >> $O/f3_probe.log 2>&1; done
for t in 0.2 0.4 0.6; do $P 8 SWEEPTT_TRIGGER_FRAC=$t >> $O/f3_probe.log 2>&1; done
for c in 1.0 1.5 2.5; do $P 8 SWEEPTT_COLCOST=$c >> $O/f3_probe.log 2>&1; done
for b in 1.5 2 3; do $P 8 SWEEPTT_BUCKET=$b >> $O/f3_probe.log 2>&1; done
cat $O/f3_probe.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()"
