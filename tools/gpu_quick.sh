#!/bin/bash
tag=${1:-r02k}
out=gpurun_out/${tag}_quick.log
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 200 python tools/profile_workload.py 10000 3 2>&1 | grep "step [12]" >> $out; }
run EMM_DONATE_AFTER=48
run EMM_LIBRARY=build_variants/lib_nodonate.so
run EMM_DONATE_AFTER=48
echo "== 4096" >> $out
EMM_DONATE_AFTER=48 timeout 200 python tools/profile_workload.py 4096 3 2>&1 | grep "step [12]" >> $out
cat $out
