#!/bin/bash
# Full evidence pass for the current build on one B200: parity tests, bench line, ncu captures, stats,
# BASELINE configs 4 / 5, sanitizers.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt; nproc >> gpurun_out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench_n1.err
tail -1 gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-files 0 > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:emm_search -c 1 -f -o gpurun_out/${tag}_search \
    python tools/profile_workload.py 2048 1 > gpurun_out/${tag}_ncu_search.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emm_prepare -c 1 -f -o gpurun_out/${tag}_prepare \
    python tools/profile_workload.py 2048 1 > gpurun_out/${tag}_ncu_prepare.log 2>&1
EMM_STATS=1 timeout 200 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1
timeout 300 python tools/stress_configs.py 296 256 > gpurun_out/${tag}_stress_configs.txt 2>&1
EMM_DONATE_AFTER=-1 timeout 300 python tools/stress_configs.py 296 0 > gpurun_out/${tag}_stress_configs_unsplit.txt 2>&1
if [ "${2:-}" = "sanitize" ]; then
timeout 800 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py 60 light > gpurun_out/${tag}_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_racecheck.log
timeout 500 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 60 light > gpurun_out/${tag}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_memcheck.log
tail -2 gpurun_out/${tag}_racecheck.log; tail -2 gpurun_out/${tag}_memcheck.log
fi
bash tools/gpu_strong.sh 1 1000000 1 4096 | tail -1 | cut -c1-200
head -6 gpurun_out/${tag}_stats.log | cut -c1-200; cat gpurun_out/${tag}_stress_configs.txt
