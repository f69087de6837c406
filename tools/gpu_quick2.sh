#!/bin/bash
tag=${1:-r02m}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "splitting or fixtures_full or synthetic or config4 or determinism" > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_quick.log; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 200 python tools/profile_workload.py 10000 3 2>&1 | grep "step [12]" >> $out; }
run EMM_DONATE_AFTER=48
cat $out
timeout 800 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py 60 light > gpurun_out/${tag}_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_racecheck.log
tail -4 gpurun_out/${tag}_racecheck.log; grep -c "hazard detected" gpurun_out/${tag}_racecheck.log
timeout 400 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 60 light > gpurun_out/${tag}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_memcheck.log
tail -3 gpurun_out/${tag}_memcheck.log
