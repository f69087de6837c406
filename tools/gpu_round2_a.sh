#!/bin/bash
# First GPU pass of round 2: parity tests, sanitizers on the final kernel, a short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
nproc >> gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py 24 > gpurun_out/r02a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 24 > gpurun_out/r02a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_memcheck.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
EMM_STATS=1 timeout 300 python tools/profile_workload.py 2048 2 > gpurun_out/r02a_stats.log 2>&1
tail -3 gpurun_out/r02a_pytest_gpu.log; tail -3 gpurun_out/r02a_racecheck.log; tail -3 gpurun_out/r02a_memcheck.log; cat gpurun_out/r02a_bench_n1.json | cut -c1-1500
