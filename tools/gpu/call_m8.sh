set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --workload config4 --gpus $n --steps 2 --warmup 1 > $O/m8d_c4_n$n.json 2> $O/m8d_c4_n$n.err
  tail -c 500 $O/m8d_c4_n$n.json; tail -2 $O/m8d_c4_n$n.err
done
