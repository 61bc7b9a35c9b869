set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python tools/stress_persistent.py 60 > $O/f4_stress.log 2>&1; tail -6 $O/f4_stress.log
timeout 900 python tools/random_parity.py 20000 250 > $O/f4_random.log 2>&1; tail -3 $O/f4_random.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/f4_bench.json 2> $O/f4_bench.err; tail -c 300 $O/f4_bench.err
timeout 900 python bench.py --workload config4 --steps 2 --warmup 1 > $O/f4_bench_c4.json 2> $O/f4_bench_c4.err; tail -c 300 $O/f4_bench_c4.err; cat $O/f4_bench_c4.json | cut -c1-900
timeout 300 python tools/probe.py 4
timeout 300 python tools/probe.py 1
