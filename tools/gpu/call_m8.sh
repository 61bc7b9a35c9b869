set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=8
nvidia-smi -L | head -8; free -g | head -2; nproc
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 > $O/m8_pytest.log; cat $O/m8_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/m8_bench_n8.json 2> $O/m8_bench_n8.err
tail -c 1200 $O/m8_bench_n8.json; tail -3 $O/m8_bench_n8.err
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extras > $O/m8_bench_n$n.json 2> $O/m8_bench_n$n.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/m8_bench_n1.json 2> $O/m8_bench_n1.err
timeout 600 python tools/config3_multi_gpu.py > $O/m8_c3_inprocess.log 2>&1; cat $O/m8_c3_inprocess.log
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 8 > $O/m8_grid.log 2>&1
$G 4 >> $O/m8_grid.log 2>&1
$G 8 SWEEPTT_NO_GLOBAL_KMIN=1 >> $O/m8_grid.log 2>&1
$G 8 SWEEPTT_BLOCK_TILES=2 >> $O/m8_grid.log 2>&1
grep -E "^\[|sweeptt\]" $O/m8_grid.log | cut -c1-400
timeout 1500 python tools/config5_bench.py --parts 1 2 4 8 --reps 1 --out $O/m8_config5.jsonl > $O/m8_config5.log 2>&1
cat $O/m8_config5.log | cut -c1-900
