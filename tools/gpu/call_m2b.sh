set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py "tests/test_gpu_full.py::test_scaled_config4_like_box_matches_reference_hashes" -m gpu -x -q 2>&1 | tail -3 > $O/m2g.log
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_dbg.so timeout 600 python tools/random_parity.py 9100 30 2>&1 | tail -2 >> $O/m2g.log
$G 1 >> $O/m2g.log 2>&1
$G 2 >> $O/m2g.log 2>&1
$G 2 8 >> $O/m2g.log 2>&1
$G 2 SWEEPTT_BUCKET=3 >> $O/m2g.log 2>&1
grep -E "^\[|sweeptt\]|passed|failed|cases" $O/m2g.log
