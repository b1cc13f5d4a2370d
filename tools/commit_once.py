"""Two device-resident commits of 2^n_log x P (the second one is the profiling target: `ncu -s <1 + launches> -c <launches>`);
prints the number of kernel launches per commit.  usage: python tools/commit_once.py [n_log=20] [P=135]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky2_gpu_b200 as p2b
n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
P = int(sys.argv[2]) if len(sys.argv) > 2 else 135
p2b.build()
ctx = p2b.Context(0)
n = 1 << n_log
vals = p2b.DeviceBuffer(ctx, P * n)
ctx.fill_synthetic(vals, P * n, 0x504C4F4E4B5932)
ctx.synchronize()
l0 = ctx.launch_count
for _ in range(2):
    b = p2b.PolynomialBatch.from_values(ctx, (vals, P, n), 3, 4)
    ctx.synchronize()
    l1 = ctx.launch_count
    print("launches per commit:", l1 - l0, "cap word 0: %016x" % int(b.cap()[0][0]))
    l0 = ctx.launch_count
    b.close()
