# ncu --set full captures of the NTT kernels (run under gpurun on one B200); raw CSV pages land in gpurun_out/
set -u
O=gpurun_out
TAG=${1:-r02b}
cap() { local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null; rm -f /tmp/$name.ncu-rep; }
cap ${TAG}_ntt_final ntt_final_pass 3 python tools/ntt_bench.py 20 135 1
