#!/bin/bash
# Second half of the evidence pass (after a kernel change): parity tests, bench line, ncu captures, config 4.
tag=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench_n1.err
tail -1 gpurun_out/${tag}_bench_n1.err; cut -c1-200 gpurun_out/${tag}_bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-files 0 > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:emm_search -c 1 -f -o gpurun_out/${tag}_search \
    python tools/profile_workload.py 2048 1 > gpurun_out/${tag}_ncu_search.log 2>&1
timeout 300 python tools/stress_configs.py 296 256 > gpurun_out/${tag}_stress_configs.txt 2>&1
EMM_DONATE_AFTER=-1 timeout 300 python tools/stress_configs.py 296 0 > gpurun_out/${tag}_stress_configs_unsplit.txt 2>&1
EMM_STATS=1 timeout 200 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1
cat gpurun_out/${tag}_stress_configs.txt gpurun_out/${tag}_stress_configs_unsplit.txt; head -6 gpurun_out/${tag}_stats.log | cut -c1-200
