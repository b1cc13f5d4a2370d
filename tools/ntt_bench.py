"""NTT pieces alone on one GPU (inverse transform, coset LDE into leaf rows): timing and a target for ncu captures.
usage: python tools/ntt_bench.py [n_log] [P] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky2_gpu_b200 as p2b

n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
P = int(sys.argv[2]) if len(sys.argv) > 2 else 135
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rate_bits = 3
p2b.build()
ctx = p2b.Context(0)
L = p2b.lib()
n = 1 << n_log
N = n << rate_bits
vals = p2b.DeviceBuffer(ctx, P * n)
ctx.fill_synthetic(vals, P * n, 0x504C4F4E4B5932)
tmp = p2b.DeviceBuffer(ctx, P * n)
leaves = p2b.DeviceBuffer(ctx, N * P)
ctx.synchronize()

def timed(fn, warm=1):
    for _ in range(warm): fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        ctx.timer_start(); fn(); ts.append(ctx.timer_stop_ms())
    return min(ts), sum(ts) / len(ts)

def ifft():
    p2b._check(L.p2b_ifft_batch(ctx.handle, vals.ptr, tmp.ptr, n_log, P))
def lde():
    p2b._check(L.p2b_lde_leaves(ctx.handle, tmp.ptr, n_log, P, rate_bits, leaves.ptr, P, 0))
a = timed(ifft)
b = timed(lde)
gb = P * n * 8 / 1e9
print("2^%d x %d: ifft min %.3f ms (%.2f GB in+out -> %.0f GB/s)   lde min %.3f ms (%.2f GB in + %.2f GB out -> %.0f GB/s algorithmic)"
      % (n_log, P, a[0], 2 * gb, 2 * gb / a[0] * 1e3, b[0], gb, 8 * gb, 9 * gb / b[0] * 1e3))
