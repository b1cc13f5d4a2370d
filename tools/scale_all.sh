#!/bin/bash
# final scaling measurement: bench.py at N = 2, 4, 8 (run under gpurun --gpus 8); JSON lines into gpurun_out/
for n in 2 4 8; do
  timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale$n.err | tail -1 > gpurun_out/scale_r01e_n$n.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/scale_r01e_n$n.json').read()); print($n, d['value'], d['e2e']['value'], d['config']['cap_word0'], d['clocks'])"
done
