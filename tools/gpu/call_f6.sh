set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_schedulers.py -m gpu -x -q 2>&1 | tail -3
