set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/f4_bench_n2.json 2> $O/f4_bench_n2.err; tail -c 900 $O/f4_bench_n2.json
