#!/bin/bash
# BASELINE config 5: commit of 2^22-2^24 rows x 135-400 columns sharded over N GPUs (run under gpurun --gpus N)
N=${1:-8}
for cfg in "22 135" "24 135" "23 400"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py \
    --gpus $N --steps 5 --warmup 3 --n-log $1 --polys $2 --no-cpu-baseline 2>gpurun_out/sweep_$1_$2.err | tail -1 > gpurun_out/sweep_c5_n${N}_$1_$2.json
  cut -c1-600 gpurun_out/sweep_c5_n${N}_$1_$2.json
done
