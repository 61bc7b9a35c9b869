set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/e2e_probe.py 111 pageable 2>&1 | grep -E "^rep" 
SWEEPTT_NO_HOST_RING=1 python tools/e2e_probe.py 111 pageable 2>&1 | grep -E "^rep"
python tools/e2e_probe.py 111 2>&1 | grep -E "^rep"
python tools/e2e_probe.py 14 pageable 2>&1 | grep -E "^rep"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/cli_e2e.py 2>&1 | tail -3
SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_dbg.so timeout 900 python tools/random_parity.py 12000 120 2>&1 | tail -3
