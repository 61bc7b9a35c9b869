set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extras > $O/m4_bench_n$n.json 2> $O/m4_bench_n$n.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/m4_bench_n1.json 2> $O/m4_bench_n1.err
timeout 300 python tools/probe.py 14
timeout 300 python tools/probe.py 28
timeout 300 python tools/probe.py 56
