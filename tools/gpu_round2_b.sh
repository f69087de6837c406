#!/bin/bash
# GPU pass: parity tests + sanitizers + short bench of the current build.  usage: gpu_round2_b.sh <tag>
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench_n1.err
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py 24 > gpurun_out/${tag}_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 24 > gpurun_out/${tag}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_memcheck.log
EMM_STATS=1 timeout 300 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1
tail -4 gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_racecheck.log; tail -2 gpurun_out/${tag}_memcheck.log; tail -2 gpurun_out/${tag}_bench_n1.err; cut -c1-900 gpurun_out/${tag}_bench_n1.json
