#!/bin/bash
# Round-2 multi-device collection (run under `gpurun --gpus N`): multi-device GPU tests, config-5 sweep, prove data path and
# the headline commit, all driven by ONE process through p2b_mgpu_*.  Usage: bash tools/collect_r02b_mgpu.sh N [with_tests]
set -u
N=${1:-8}
O=gpurun_out
mkdir -p $O
if [ "${2:-}" = "tests" ]; then
  timeout 600 python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_sharded.py tests/test_gpu_configs.py -m gpu -q > $O/r02b_mgpu_tests_n$N.txt 2>&1
  tail -2 $O/r02b_mgpu_tests_n$N.txt
fi
timeout 900 python tools/sweep_c5_mgpu.py $N 2 > $O/r02b_sweep_c5_n$N.jsonl 2> $O/r02b_sweep_c5_n$N.err
cut -c1-260 $O/r02b_sweep_c5_n$N.jsonl
timeout 300 python tools/mgpu_prove_bench.py $N recursion 20 3 2>&1 | tail -1 > $O/r02b_mgpu_prove_recursion_n$N.json
timeout 300 python tools/mgpu_prove_bench.py $N ecc 17 3 2>&1 | tail -1 > $O/r02b_mgpu_prove_ecc_n$N.json
timeout 300 python tools/mgpu_bench.py $N 20 135 5 2>&1 | tail -1 > $O/r02b_mgpu_commit_n$N.json
cat $O/r02b_mgpu_prove_recursion_n$N.json $O/r02b_mgpu_prove_ecc_n$N.json $O/r02b_mgpu_commit_n$N.json
