"""Quotient-polynomial evaluation at BASELINE config 4 scale (U32-gate-heavy circuit, ~2^17 rows, wide_ecc_config 234 wires):
times p2b_quotient_polys on synthetic committed matrices (random field elements: the kernel's work does not depend on
whether the constraints hold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plonky2_gpu_b200 as p2b

n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 17
num_wires, num_routed, num_constants = 234, 80, 8
ctx = p2b.Context(0)
n = 1 << n_log
def commit(cols, seed):
    d = p2b.DeviceBuffer(ctx, cols * n)
    ctx.fill_synthetic(d, cols * n, seed)
    return p2b.PolynomialBatch.from_values(ctx, (d, cols, n), 3, 4)
bw, bz, bc = commit(num_wires, 1), commit(20, 2), commit(num_constants + num_routed, 3)
# an ed25519-style gate mix (cuda/plonky2_gpu_impl.cuh:600-685 lists 25 instances of these types)
G = p2b
gates = [(G.GATE_NOOP, ()), (G.GATE_CONSTANT, (2,)), (G.GATE_PUBLIC_INPUT, ()), (G.GATE_ARITHMETIC, (20,)), (G.GATE_BASE_SUM, (63, 2)),
         (G.GATE_BASE_SUM, (32, 2)), (G.GATE_RANDOM_ACCESS, (4, 4, 2)), (G.GATE_RANDOM_ACCESS, (2, 13, 2)), (G.GATE_U32_ARITHMETIC, (6,)),
         (G.GATE_U32_ADD_MANY, (3, 9)), (G.GATE_U32_ADD_MANY, (5, 8)), (G.GATE_U32_RANGE_CHECK, (8,)), (G.GATE_U32_SUBTRACTION, (11,)),
         (G.GATE_COMPARISON, (32, 16)), (G.GATE_COMPARISON, (8, 4)), (G.GATE_POSEIDON, ())]
groups = [(0, 4), (4, 8), (8, 12), (12, 15), (15, 16)]
sel = [0] * 4 + [1] * 4 + [2] * 4 + [3] * 3 + [4]
circ = p2b.Circuit(gates, sel, groups, num_wires, num_routed, num_constants, [pow(7, j, p2b.ORDER) for j in range(num_routed)], n_log)
rng = np.random.default_rng(0)
rnd = lambda k: [int(x) for x in rng.integers(0, p2b.ORDER, size=k, dtype=np.uint64)]
pih, betas, gammas, alphas = rnd(4), rnd(2), rnd(2), rnd(2)
import ctypes as C
arr = lambda x, k: (C.c_uint64 * k)(*x)
size = circ.lde_size
dv, dc = p2b.DeviceBuffer(ctx, 2 * size), p2b.DeviceBuffer(ctx, 2 * size)
def run():
    p2b._check(p2b.lib().p2b_quotient_polys(ctx.handle, C.byref(circ.struct), bw.handle, bz.handle, bc.handle, arr(pih, 4), arr(betas, 2),
                                            arr(gammas, 2), arr(alphas, 2), dv.ptr, dc.ptr))
run(); ctx.synchronize()
ts = []
for _ in range(3):
    ctx.timer_start(); run(); ts.append(ctx.timer_stop_ms())
print("quotient 2^%d rows x %d wires, %d gates, %d LDE points: %.2f ms (min of 3)" % (n_log, num_wires, len(gates), size, min(ts)))
