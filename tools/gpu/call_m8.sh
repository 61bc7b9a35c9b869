set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
nvidia-smi -L | head -8; free -g | head -2; nproc
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 > $O/m8b_pytest.log; cat $O/m8b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/m8b_bench_n8.json 2> $O/m8b_bench_n8.err
tail -c 900 $O/m8b_bench_n8.json; tail -3 $O/m8b_bench_n8.err
timeout 600 python tools/config3_multi_gpu.py > $O/m8b_c3_inprocess.log 2>&1; cat $O/m8b_c3_inprocess.log
G="timeout 300 python tools/grid_probe.py 1201 1201 251"
$G 8 > $O/m8b_grid.log 2>&1
$G 8 8 >> $O/m8b_grid.log 2>&1
$G 8 SWEEPTT_BUCKET=3 >> $O/m8b_grid.log 2>&1
$G 8 SWEEPTT_NO_GLOBAL_KMIN=1 >> $O/m8b_grid.log 2>&1
$G 8 SWEEPTT_BLOCK_TILES=5 >> $O/m8b_grid.log 2>&1
$G 4 >> $O/m8b_grid.log 2>&1
grep -E "^\[|sweeptt\]" $O/m8b_grid.log | cut -c1-330
timeout 1500 python tools/config5_bench.py --parts 1 2 4 8 --reps 1 --out $O/m8b_config5.jsonl > $O/m8b_config5.log 2>&1
cat $O/m8b_config5.log | cut -c1-700
