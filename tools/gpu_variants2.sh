#!/bin/bash
# A/B at bench size (10 000 structures): what does the pair-splitting machinery cost on the headline workload?
tag=${1:-r02h}
out=gpurun_out/${tag}_variants.log
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 200 python tools/profile_workload.py 10000 3 2>&1 | grep "step [12]" >> $out; }
run EMM_DONATE_AFTER=48
run EMM_LIBRARY=build_variants/lib_nodonate.so
run EMM_DONATE_AFTER=16
run EMM_DONATE_AFTER=200
echo "== 4096 structures" >> $out
EMM_DONATE_AFTER=48 timeout 200 python tools/profile_workload.py 4096 3 2>&1 | grep "step [12]" >> $out
for d in 48; do
  echo "== config4 EMM_DONATE_AFTER=$d" >> $out
  EMM_DONATE_AFTER=$d timeout 200 python tools/stress_configs.py 296 0 >> $out 2>&1
done
cat $out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
EMM_STATS=1 timeout 200 python tools/profile_workload.py 2048 2 > gpurun_out/${tag}_stats.log 2>&1; head -8 gpurun_out/${tag}_stats.log | cut -c1-200
