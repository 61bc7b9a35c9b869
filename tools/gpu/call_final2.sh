set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/f2_pytest.log; cat $O/f2_pytest.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/f2_bench.json 2> $O/f2_bench.err; tail -c 300 $O/f2_bench.err
timeout 120 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > $O/f2_ref.json 2> $O/f2_ref.err; tail -c 300 $O/f2_ref.err
