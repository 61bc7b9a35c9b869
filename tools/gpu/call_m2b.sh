set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/config3_multi_gpu.py
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/m2i_bench.json 2> $O/m2i_bench.err; tail -c 700 $O/m2i_bench.json
