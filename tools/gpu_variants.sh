#!/bin/bash
# A/B of kernel variants and schedule knobs on one box: kernel ms for 4096 structures x full library.
tag=${1:-r02e}
out=gpurun_out/${tag}_variants.log
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 150 python tools/profile_workload.py 4096 3 2>&1 | grep "step [12]" >> $out; }
run EMM_DONATE_AFTER=-1
run EMM_DONATE_AFTER=48
run EMM_DONATE_AFTER=12
run EMM_DONATE_AFTER=200
run EMM_TWO_PHASE=0 EMM_DONATE_AFTER=48
run EMM_TWO_PHASE=0 EMM_DONATE_AFTER=12
run EMM_LIBRARY=build_variants/lib_t704.so EMM_DONATE_AFTER=48
run EMM_LIBRARY=build_variants/lib_t704.so EMM_DONATE_AFTER=-1
run EMM_LIBRARY=build_variants/lib_t640.so EMM_DONATE_AFTER=48
run EMM_LIBRARY=build_variants/lib_t640.so EMM_DONATE_AFTER=-1
for d in -1 48; do
  echo "== config4 EMM_DONATE_AFTER=$d" >> $out
  EMM_DONATE_AFTER=$d timeout 200 python tools/stress_configs.py 296 0 >> $out 2>&1
done
echo "== config5 default" >> $out
timeout 200 python tools/stress_configs.py 0 256 >> $out 2>&1
EMM_STATS=1 timeout 200 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1
cat $out; head -8 gpurun_out/${tag}_stats.log | cut -c1-200
