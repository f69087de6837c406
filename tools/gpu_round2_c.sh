#!/bin/bash
# Guarded GPU pass for a risky kernel change: short timeouts everywhere, stops at the first failure.
# usage: gpu_round2_c.sh <tag>
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 300 python tools/sanitizer_workload.py 24 > gpurun_out/${tag}_smoke.log 2>&1; rc=$?; echo "workload rc=$rc" >> gpurun_out/${tag}_smoke.log
if [ $rc -ne 0 ]; then
  tail -5 gpurun_out/${tag}_smoke.log
  timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 24 > gpurun_out/${tag}_memcheck.log 2>&1
  grep -m 12 -A12 "Invalid\|ERROR SUMMARY\|hang" gpurun_out/${tag}_memcheck.log | head -60
  exit 1
fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${tag}_pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench_n1.err
tail -2 gpurun_out/${tag}_bench_n1.err; cut -c1-1200 gpurun_out/${tag}_bench_n1.json
for env in "EMM_DONATE_AFTER=48"; do
  echo "== $env" >> gpurun_out/${tag}_variants.log
  env $env EMM_STATS=0 timeout 120 python tools/profile_workload.py 4096 3 >> gpurun_out/${tag}_variants.log 2>&1
done
EMM_STATS=1 timeout 200 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1
timeout 300 python tools/stress_configs.py 296 256 > gpurun_out/${tag}_stress.log 2>&1
cat gpurun_out/${tag}_variants.log | grep "==\|step 2"; tail -12 gpurun_out/${tag}_stats.log | head -6; cat gpurun_out/${tag}_stress.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py 24 > gpurun_out/${tag}_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_racecheck.log
tail -2 gpurun_out/${tag}_racecheck.log
