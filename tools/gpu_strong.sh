#!/bin/bash
# BASELINE config 3: one list of <total> structures over N GPUs, merge timed.  usage: gpu_strong.sh N [total] [steps] [chunk]
N=${1:-1}; total=${2:-1000000}; steps=${3:-1}; chunk=${4:-4096}
mkdir -p gpurun_out
out=gpurun_out/r02_strong_n${N}.json
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --scaling strong --gpus 1 --steps $steps --warmup 1 --total $total --chunk $chunk > $out 2> gpurun_out/r02_strong_n${N}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --scaling strong --gpus $N --steps $steps --warmup 1 --total $total --chunk $chunk > $out 2> gpurun_out/r02_strong_n${N}.err
fi
echo "rc=$?"; tail -3 gpurun_out/r02_strong_n${N}.err; cut -c1-1500 $out
