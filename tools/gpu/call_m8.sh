set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/m8c_bench_n8.json 2> $O/m8c_bench_n8.err
tail -c 600 $O/m8c_bench_n8.json; tail -3 $O/m8c_bench_n8.err
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extras > $O/m8c_bench_n$n.json 2> $O/m8c_bench_n$n.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/m8c_bench_n1.json 2> $O/m8c_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/m8c_ref_n8.json 2> $O/m8c_ref_n8.err
cat $O/m8c_ref_n8.json | cut -c1-400
timeout 600 python tools/config3_multi_gpu.py > $O/m8c_c3_inprocess.log 2>&1; cat $O/m8c_c3_inprocess.log
