# usage: bash tools/gpu/call_multi.sh N   (N GPUs on the box)
set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
O=gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -15 > $O/m${N}_pytest.log
cat $O/m${N}_pytest.log
timeout 600 python tools/config5_bench.py --dims 1201 1201 251 --seed 11 --parts 1 $N --reps 2 --out $O/m${N}_c4grid.jsonl > $O/m${N}_c4grid.log 2>&1
tail -4 $O/m${N}_c4grid.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/m${N}_bench.json 2> $O/m${N}_bench.err
tail -c 1500 $O/m${N}_bench.json; tail -5 $O/m${N}_bench.err
if [ "$N" -ge 4 ]; then
  for n in 1 2 4 8; do
    if [ $n -le $N ] && [ $n -ne $N ]; then
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $O/m${N}_bench_n$n.json 2> $O/m${N}_bench_n$n.err
    fi
  done
  timeout 600 python tools/config3_multi_gpu.py > $O/m${N}_c3_inprocess.log 2>&1
  cat $O/m${N}_c3_inprocess.log
  timeout 1500 python tools/config5_bench.py --parts 1 2 4 8 --reps 1 --out $O/m${N}_config5.jsonl > $O/m${N}_config5.log 2>&1
  tail -6 $O/m${N}_config5.log
fi
