set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/cli_e2e.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli or output" 2>&1 | tail -3
