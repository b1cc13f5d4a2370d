#!/bin/bash
# Final round-2 collection on one B200 (run under gpurun): GPU tests, bench line, launch list, ncu captures reduced to CSV,
# prove pipelines.  Outputs: gpurun_out/r02c_*
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r02c_gputests.txt 2>&1; tail -2 $O/r02c_gputests.txt
python tools/ntt_bench.py 20 135 > $O/r02c_ntt_bench.txt 2>&1; cat $O/r02c_ntt_bench.txt
python bench.py --steps 10 --warmup 3 > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r02c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > /dev/null 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r02c_hash_leaves_bench hash_leaves 16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda
cap r02c_intt_final intt_final_pass 1 python tools/ntt_bench.py 20 135 1
python tools/prove_pipeline.py ecc 17 3 > $O/r02c_prove_ecc.txt 2>&1
python tools/prove_pipeline.py recursion 16 3 > $O/r02c_prove_rec16.txt 2>&1
python tools/prove_pipeline.py recursion 20 3 > $O/r02c_prove_rec20.txt 2>&1
python bench.py --workload prove-recursion --n-log 20 --steps 3 --warmup 3 > $O/r02c_bench_prove_recursion.json 2>/dev/null
python bench.py --workload prove-ecc --steps 3 --warmup 3 > $O/r02c_bench_prove_ecc.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.txt 2>&1; tail -1 $O/r02c_smoke.txt
cut -c1-400 $O/r02c_bench_n1.json
