"""Scratch timing of the commit and its pieces on one GPU (not the judged bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import plonky2_gpu_b200 as p2b

n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
P = int(sys.argv[2]) if len(sys.argv) > 2 else 135
rate_bits, cap_height = 3, 4
p2b.build()
ctx = p2b.Context(0)
L = p2b.lib()
n = 1 << n_log
vals = p2b.DeviceBuffer(ctx, P * n)
ctx.fill_synthetic(vals, P * n, 0x504C4F4E4B5932)
ctx.synchronize()

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        ctx.timer_start(); fn(); ts.append(ctx.timer_stop_ms())
    return min(ts), sum(ts) / len(ts)

def commit():
    b = p2b.PolynomialBatch.from_values(ctx, (vals, P, n), rate_bits, cap_height)
    b.close()
print("commit from_values 2^%d x %d: min %.2f ms avg %.2f ms" % ((n_log, P) + timed(commit)))

# pieces
tmp = p2b.DeviceBuffer(ctx, P * n)
def ifft():
    p2b._check(L.p2b_ifft_batch(ctx.handle, vals.ptr, tmp.ptr, n_log, P))
print("ifft: min %.2f ms avg %.2f" % timed(ifft))
N = n << rate_bits
leaves = p2b.DeviceBuffer(ctx, N * P)
def lde():
    p2b._check(L.p2b_lde_leaves(ctx.handle, tmp.ptr, n_log, P, rate_bits, leaves.ptr, P, 0))
print("lde: min %.2f ms avg %.2f" % timed(lde))
dig = p2b.DeviceBuffer(ctx, 2 * N * 4)
cap = p2b.DeviceBuffer(ctx, 16 * 4)
def merkle():
    p2b._check(L.p2b_merkle_tree(ctx.handle, leaves.ptr, N, P, P, 1, cap_height, dig.ptr, cap.ptr))
print("merkle: min %.2f ms avg %.2f" % timed(merkle))
# raw permutation throughput
cnt = 148 * 128 * 64
st = p2b.DeviceBuffer(ctx, cnt * 12)
ctx.fill_synthetic(st, cnt * 12, 1)
def perm():
    p2b._check(L.p2b_poseidon_permute(ctx.handle, st.ptr, cnt))
mn, av = timed(perm, reps=5)
print("permute x%d: min %.3f ms -> %.1f Mperm/s" % (cnt, mn, cnt / mn / 1e3))
print("launches", ctx.launch_count)
