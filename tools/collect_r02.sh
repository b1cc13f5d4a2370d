#!/bin/bash
# Round-2 measurement collection (run under gpurun on one B200): bench line, launch list, ncu captures reduced to CSV
# (the .ncu-rep files are too large to bring back), prove pipelines.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > /dev/null 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r02_hash_leaves_bench hash_leaves 16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda
cap r02c_quotient_recursion quotient_values 0 python tools/prove_pipeline.py recursion 17 1
cap r02c_quotient_ecc quotient_values 0 python tools/prove_pipeline.py ecc 17 1
python tools/prove_pipeline.py ecc 17 3 > $O/r02_prove_ecc.txt 2>&1
python tools/prove_pipeline.py recursion 16 3 > $O/r02_prove_rec16.txt 2>&1
python tools/prove_pipeline.py recursion 20 3 > $O/r02_prove_rec20.txt 2>&1
python tools/fri_bench.py > $O/r02_fri.txt 2>&1
python bench.py --workload prove-recursion --n-log 20 --steps 3 --warmup 3 > $O/r02_bench_prove_recursion.json 2>/dev/null
python bench.py --workload prove-ecc --steps 3 --warmup 3 > $O/r02_bench_prove_ecc.json 2>/dev/null
du -sh $O
