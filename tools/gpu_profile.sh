#!/bin/bash
# ncu evidence for the current build (one GPU): launch list of a short bench run, --set full captures of the
# search and prepare kernels on 2048 structures x full library, and instruction counts with / without splitting.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-files 0 > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emm_search -c 1 -f -o gpurun_out/${tag}_search \
    python tools/profile_workload.py 2048 1 > gpurun_out/${tag}_ncu_search.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emm_prepare -c 1 -f -o gpurun_out/${tag}_prepare \
    python tools/profile_workload.py 2048 1 > gpurun_out/${tag}_ncu_prepare.log 2>&1
for v in base nodonate; do
  lib=""; [ $v = nodonate ] && lib="EMM_LIBRARY=build_variants/lib_nodonate.so"
  timeout 300 env $lib ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
      -k regex:emm_search -c 1 --csv --log-file gpurun_out/${tag}_inst_$v.csv python tools/profile_workload.py 2048 1 > /dev/null 2>&1
  grep -h "emm_search" gpurun_out/${tag}_inst_$v.csv | cut -d, -f 5,12- | head -4
done
ls -la gpurun_out/${tag}_*.ncu-rep; tail -2 gpurun_out/${tag}_ncu_search.log
