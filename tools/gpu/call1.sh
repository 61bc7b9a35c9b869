set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
for w in 0 4 6; do SWEEPTT_WAVE=$w timeout 300 python tools/probe.py 111 SWEEPTT_WAVE=$w >> gpurun_out/c1_probe.log 2>&1; done
timeout 300 python tools/probe.py 111 >> gpurun_out/c1_probe.log 2>&1
timeout 300 python tools/probe.py 8 >> gpurun_out/c1_probe.log 2>&1
timeout 300 python tools/probe.py 4 >> gpurun_out/c1_probe.log 2>&1
timeout 300 python tools/probe.py 1 >> gpurun_out/c1_probe.log 2>&1
export SWEEPTT_LIB=$GRAFT_REPO_ROOT/uoparallel_seismic_project_b200/lib_exp/libsweeptt_tx4.so
timeout 300 python tools/probe.py 4 TX=4 >> gpurun_out/c1_probe.log 2>&1
timeout 300 python tools/probe.py 8 TX=4 >> gpurun_out/c1_probe.log 2>&1
timeout 300 python tools/probe.py 111 TX=4 SWEEPTT_WAVE=0 >> gpurun_out/c1_probe.log 2>&1
unset SWEEPTT_LIB
timeout 1200 python tools/legacy_gpu.py > gpurun_out/c1_legacy.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c1_ref.json 2> gpurun_out/c1_ref.err
cat gpurun_out/c1_pytest.log gpurun_out/c1_probe.log; tail -3 gpurun_out/c1_legacy.log; tail -c 600 gpurun_out/c1_bench.err
