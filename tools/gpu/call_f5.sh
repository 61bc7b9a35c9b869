set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 > $O/f5_bench.json 2> $O/f5_bench.err; tail -c 200 $O/f5_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/f5_ref.json 2> $O/f5_ref.err
timeout 300 python tools/probe.py 10 SWEEPTT_PERSIST_MAX_KEYS=21000
