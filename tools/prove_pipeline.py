"""The data path of prove() (plonky2/src/plonk/prover.rs:239-700 `my_prove`) end to end on the device, on a synthetic
witness: everything between "the witness matrix exists" and "the proof's field elements exist", i.e.

  commit wires -> Z / partial products -> commit them -> quotient values + coefficients -> commit the quotient chunks ->
  openings at zeta, g*zeta -> FRI opening proof (commit phase, PoW, query rounds)

with the transcript steps between the calls done by a stand-in (fixed challenges: the work does not depend on their values).
Not included (control plane, out of scope): circuit building, witness generation, the preprocessed constants_sigmas
commit (done once per circuit; built here outside the timed region like the reference's CircuitData).

  python tools/prove_pipeline.py ecc 17        # BASELINE config 4 shape: 2^17 rows, wide_ecc_config (234 wires), U32-heavy gates
  python tools/prove_pipeline.py recursion 16  # BASELINE config 3 shape: standard_recursion_config (135 wires), recursion gates
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plonky2_gpu_b200 as p2b

kind = sys.argv[1] if len(sys.argv) > 1 else "ecc"
n_log = int(sys.argv[2]) if len(sys.argv) > 2 else 17
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
G = p2b
if kind == "ecc":
    num_wires, num_routed, num_gate_consts = 234, 80, 2
    gates = [(G.GATE_NOOP, ()), (G.GATE_CONSTANT, (2,)), (G.GATE_PUBLIC_INPUT, ()), (G.GATE_ARITHMETIC, (20,)), (G.GATE_BASE_SUM, (63, 2)),
             (G.GATE_BASE_SUM, (32, 2)), (G.GATE_RANDOM_ACCESS, (4, 4, 2)), (G.GATE_RANDOM_ACCESS, (2, 13, 2)), (G.GATE_U32_ARITHMETIC, (6,)),
             (G.GATE_U32_ADD_MANY, (3, 9)), (G.GATE_U32_ADD_MANY, (5, 8)), (G.GATE_U32_RANGE_CHECK, (8,)), (G.GATE_U32_SUBTRACTION, (11,)),
             (G.GATE_COMPARISON, (32, 16)), (G.GATE_COMPARISON, (8, 4)), (G.GATE_POSEIDON, ())]
    groups = [(0, 4), (4, 8), (8, 12), (12, 15), (15, 16)]
    sel = [0] * 4 + [1] * 4 + [2] * 4 + [3] * 3 + [4]
else:
    num_wires, num_routed, num_gate_consts = 135, 80, 2
    gates = [(G.GATE_NOOP, ()), (G.GATE_CONSTANT, (2,)), (G.GATE_PUBLIC_INPUT, ()), (G.GATE_ARITHMETIC, (20,)), (G.GATE_ARITHMETIC_EXTENSION, (10,)),
             (G.GATE_MUL_EXTENSION, (13,)), (G.GATE_REDUCING, (43,)), (G.GATE_REDUCING_EXTENSION, (32,)), (G.GATE_BASE_SUM, (63, 2)),
             (G.GATE_RANDOM_ACCESS, (4, 4, 2)), (G.GATE_EXPONENTIATION, (66,)), (G.GATE_POSEIDON_MDS, ()),
             (G.GATE_LOW_DEGREE_INTERPOLATION, (4,)), (G.GATE_POSEIDON, ())]
    groups = [(0, 5), (5, 9), (9, 12), (12, 13), (13, 14)]
    sel = [0] * 5 + [1] * 4 + [2] * 3 + [3] + [4]
rate_bits, cap_height, nc, qdf, pow_bits, queries = 3, 4, 2, 8, 16, 28
num_constants = len(groups) + num_gate_consts
n = 1 << n_log
K = -(-num_routed // qdf)
arity, d = [], n_log
while d > 5 and d + rate_bits - 4 >= cap_height:
    arity.append(4)
    d -= 4

ctx = p2b.Context(0)
L = p2b.lib()
k_is = [pow(7, j, p2b.ORDER) for j in range(num_routed)]
circ = p2b.Circuit(gates, sel, groups, num_wires, num_routed, num_constants, k_is, n_log, rate_bits, nc, qdf)
rng = np.random.default_rng(0)
rnd = lambda k: [int(x) for x in rng.integers(0, p2b.ORDER, size=k, dtype=np.uint64)]
pih, betas, gammas, alphas = rnd(4), rnd(nc), rnd(nc), rnd(nc)
arr = lambda x: (C.c_uint64 * len(x))(*x)


def synth(cols, seed):
    dbuf = p2b.DeviceBuffer(ctx, cols * n)
    ctx.fill_synthetic(dbuf, cols * n, seed)
    return dbuf


# per-circuit data (outside the timed region): constants_sigmas values and their commitment
d_cs = synth(num_constants + num_routed, 3)
b_cs = p2b.PolynomialBatch.from_values(ctx, (d_cs, num_constants + num_routed, n), rate_bits, cap_height)
d_sigma = p2b.DeviceBuffer(ctx, num_routed * n)   # sigma values = the last num_routed columns of constants_sigmas
sig_host = np.empty(num_routed * n, dtype=np.uint64)
L.p2b_memcpy_d2h(ctx.handle, sig_host.ctypes.data, C.c_void_p(d_cs.ptr + 8 * num_constants * n), 8 * num_routed * n)
L.p2b_memcpy_h2d(ctx.handle, d_sigma.ptr, sig_host.ctypes.data, 8 * num_routed * n)
d_wires = synth(num_wires, 1)                     # the witness (random: every kernel's work is data-independent)
ctx.synchronize()

zeta = (0x123456789abcdef, 0xfedcba987654321)
g = pow(1753635133440165772, 1 << (32 - n_log), p2b.ORDER)
zeta_next = (zeta[0] * g % p2b.ORDER, zeta[1] * g % p2b.ORDER)
size = circ.lde_size


def stage(name, fn, times):
    ctx.timer_start()
    r = fn()
    times.setdefault(name, []).append(ctx.timer_stop_ms())
    return r


def prove(times):
    t_all = time.perf_counter()
    b_w = stage("commit wires", lambda: p2b.PolynomialBatch.from_values(ctx, (d_wires, num_wires, n), rate_bits, cap_height), times)
    zs, shape = stage("Z + partial products", lambda: p2b.partial_products_and_zs(ctx, (d_wires, num_wires, n), (d_sigma, num_routed, n),
                                                                                  k_is, betas, gammas, qdf), times)
    b_z = stage("commit Z/pp", lambda: p2b.PolynomialBatch.from_values(ctx, (zs, shape[0], n), rate_bits, cap_height), times)
    dv, dc = p2b.DeviceBuffer(ctx, nc * size), p2b.DeviceBuffer(ctx, nc * size)
    stage("quotient polys", lambda: p2b._check(L.p2b_quotient_polys(ctx.handle, C.byref(circ.struct), b_w.handle, b_z.handle, b_cs.handle,
                                                                     arr(pih), arr(betas), arr(gammas), arr(alphas), dv.ptr, dc.ptr)), times)
    # quotient_poly.chunks(degree) (prover.rs:151-166): [nc][8n] coefficients are already [nc*8][n] chunk-major
    b_q = stage("commit quotient chunks", lambda: p2b.PolynomialBatch.from_coeffs(ctx, (dc, nc * qdf, n), rate_bits, cap_height), times)
    oracles = [b_cs, b_w, b_z, b_q]

    def openings():
        for o in oracles:
            p2b.eval_openings(ctx, o, zeta)
        p2b.eval_openings(ctx, b_z, zeta_next)
    stage("openings", openings, times)
    polys = (num_constants + num_routed, num_wires, nc * K, nc * qdf)
    all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
    ch = p2b.Challenger(list(range(1, 13)), [5, 6, 7])
    pr = stage("FRI prove_openings", lambda: p2b.fri_prove_openings(ctx, oracles, [(zeta, all_polys), (zeta_next, [(2, p) for p in range(nc)])], ch,
                                                                    n_log, rate_bits, cap_height, pow_bits, queries, arity), times)
    ctx.synchronize()
    wall = (time.perf_counter() - t_all) * 1e3
    pr.close()
    for b in (b_w, b_z, b_q):
        b.close()
    return wall


prove({})
times, walls = {}, []
for _ in range(reps):
    walls.append(prove(times))
print("prove() data path, %s shape: 2^%d rows x %d wires, %d gates, rate %d, %d FRI reductions" % (kind, n_log, num_wires, len(gates), rate_bits, len(arity)))
total = 0.0
for k, v in times.items():
    print("  %-26s %9.2f ms" % (k, min(v)))
    total += min(v)
print("  %-26s %9.2f ms (sum of stage minima; host wall per proof %.2f ms, min of %d)" % ("total", total, min(walls), reps))
