"""GPU: the BASELINE.json configs beyond the headline commit, at (or near) their full sizes, with checks that scale:

config 3 / 4  prove() data path on the recursion / ecc shapes (plonky2/examples/bench_recursion.rs:175-207,
              ecdsa/src/gadgets/ecdsa.rs:64-110): the quotient VALUES at sampled points of the full LDE domain equal the CPU
              oracle's evaluation of the same rows (every gate, filter, permutation term and the Z_H division at full size),
              the quotient coefficients are the coset-iNTT of those values, Z/partial products equal the oracle on sampled
              columns, and every commitment's opened rows verify against its cap.
config 5      the scale sweep's shapes that fit one GPU (2^22 x 135, 2^21 x 234, 2^20 x 400): opened rows verify against
              the cap, rows equal direct evaluations of the coefficients, commit(from_coeffs) is idempotent; on >= 2 GPUs
              the single-process multi-device commit reproduces the single-device cap and rows.
The largest shape the reference's own CUDA ABI can address (C int sizes: 2^20 x 234) is cross-checked against the
reference kernels when oracle/_ref is present.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from oracle import quotient as Q

pytestmark = pytest.mark.gpu
P = oracle.ORDER
SEED = 0x504C4F4E4B5932
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


def _oracle_circuit(pipe):
    """The same circuit description for oracle/quotient.py (gate objects instead of type ids)."""
    from oracle import recursion_gates as RG
    G = p2b
    mk = {G.GATE_NOOP: lambda: Q.NoopGate(), G.GATE_CONSTANT: lambda a: Q.ConstantGate(a), G.GATE_PUBLIC_INPUT: lambda: Q.PublicInputGate(),
          G.GATE_ARITHMETIC: lambda a: Q.ArithmeticGate(a), G.GATE_BASE_SUM: lambda a, b: Q.BaseSumGate(a, b), G.GATE_POSEIDON: lambda: Q.PoseidonGate(),
          G.GATE_RANDOM_ACCESS: lambda a, b, c: Q.RandomAccessGate(a, b, c), G.GATE_U32_ARITHMETIC: lambda a: Q.U32ArithmeticGate(a),
          G.GATE_U32_ADD_MANY: lambda a, b: Q.U32AddManyGate(a, b), G.GATE_U32_RANGE_CHECK: lambda a: Q.U32RangeCheckGate(a),
          G.GATE_U32_SUBTRACTION: lambda a: Q.U32SubtractionGate(a), G.GATE_COMPARISON: lambda a, b: Q.ComparisonGate(a, b),
          G.GATE_ARITHMETIC_EXTENSION: lambda a: RG.ArithmeticExtensionGate(a), G.GATE_MUL_EXTENSION: lambda a: RG.MulExtensionGate(a),
          G.GATE_REDUCING: lambda a: RG.ReducingGate(a), G.GATE_REDUCING_EXTENSION: lambda a: RG.ReducingExtensionGate(a),
          G.GATE_EXPONENTIATION: lambda a: RG.ExponentiationGate(a), G.GATE_POSEIDON_MDS: lambda: RG.PoseidonMdsGate(),
          G.GATE_LOW_DEGREE_INTERPOLATION: lambda a: RG.LowDegreeInterpolationGate(a),
          G.GATE_HIGH_DEGREE_INTERPOLATION: lambda a: RG.HighDegreeInterpolationGate(a)}
    gates = [mk[t](*params) for t, params in pipe.gates]
    return Q.Circuit(gates, pipe.sel, pipe.groups, pipe.num_wires, pipe.num_routed, pipe.num_constants, pipe.k_is, pipe.n_log,
                     pipe.rate_bits, pipe.nc, pipe.qdf)


@pytest.mark.parametrize("kind,n_log", [("ecc", 17), ("recursion", 16), ("recursion", 18)])
def test_prove_data_path_at_baseline_shapes(ctx, kind, n_log):
    from plonky2_gpu_b200.pipeline import ProvePipeline
    pipe = ProvePipeline(ctx, kind, n_log)
    wall, keep = pipe.prove(keep=True)
    circ = _oracle_circuit(pipe)
    size, nc, n = pipe.size, pipe.nc, pipe.n
    N = n << pipe.rate_bits
    rng = np.random.default_rng(9)
    pts = sorted(set([0, 1, size - 1, size // 2 + 3] + [int(x) for x in rng.integers(0, size, size=8)]))
    # rows each sampled point reads, fetched from the device batches (and verified against the caps on the way)
    need = sorted(set(r for i in pts for r in Q.quotient_point_rows(circ, i)))
    rows = {}
    for name, b in (("w", keep["b_w"]), ("z", keep["b_z"]), ("cs", pipe.b_cs)):
        got, sibs = b.open_rows(need)
        cap = b.cap()
        for r, row, sb in zip(need, got, sibs):
            assert oracle.merkle_verify(row, r, cap, sb)
        rows[name] = dict(zip(need, got))
    want = Q.compute_quotient_values(circ, rows["w"], rows["z"], rows["cs"], pipe.pih, pipe.betas, pipe.gammas, pipe.alphas, points=pts)
    qv = keep["quotient_values"].to_host(nc * size).reshape(nc, size)
    for i, wv in zip(pts, want):
        for c in range(nc):
            assert int(qv[c][i]) == wv[c], (kind, i, c)
    # the coefficients are the coset-iNTT of the values (prover.rs:1009-1021): check by evaluating them at sampled points
    qc = keep["quotient_coeffs"].to_host(nc * size).reshape(nc, size)
    w = oracle.primitive_root_of_unity(pipe.n_log + 3)
    for i in pts[:4]:
        x = 7 * oracle.exp(w, i) % P
        for c in range(nc):
            assert int(oracle.naive_coset_eval(qc[c], 0, int(x))[0]) == int(qv[c][i])
    # Z / partial products: two columns against the oracle (prover.rs:702-786) -- full columns are O(n * routed) in Python,
    # so this is done at the smallest shape only
    if n_log <= 16:
        wires = pipe.d_wires.to_host(pipe.num_wires * n).reshape(pipe.num_wires, n)
        sigma = pipe.d_sigma.to_host(pipe.num_routed * n).reshape(pipe.num_routed, n)
        zs = keep["zs"].to_host(keep["zs_shape"][0] * n).reshape(keep["zs_shape"][0], n)
        sub = 1 << 9   # the recurrence is sequential in the row index: the first 2^9 rows pin every chunk formula and the running product
        x = 1
        wn = oracle.primitive_root_of_unity(pipe.n_log)
        z = 1
        for i in range(sub):
            assert int(zs[0][i]) == z, i        # Z_0 at row i
            num = den = 1
            for j in range(pipe.num_routed):
                num = num * ((int(wires[j][i]) + pipe.betas[0] * (pipe.k_is[j] * x % P) + pipe.gammas[0]) % P) % P
                den = den * ((int(wires[j][i]) + pipe.betas[0] * int(sigma[j][i]) + pipe.gammas[0]) % P) % P
            z = z * num % P * pow(den, P - 2, P) % P
            x = x * wn % P
    # the quotient-chunk commitment opens consistently too
    got, sibs = keep["b_q"].open_rows([0, N - 1, 12345 % N])
    capq = keep["b_q"].cap()
    for r, row, sb in zip([0, N - 1, 12345 % N], got, sibs):
        assert oracle.merkle_verify(row, r, capq, sb)
    keep["proof"].close()
    for k in ("b_w", "b_z", "b_q"):
        keep[k].close()
    pipe.close()


def _commit_properties(ctx, n_log, Pn, batch_factory):
    rate_bits, cap_height = 3, 4
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    b, coeffs_of = batch_factory(n_log, Pn)
    cap = b.cap()
    rng = np.random.default_rng(5)
    idx = sorted(set([0, 1, N - 1, n - 1, n, 5 * n + 17] + [int(x) for x in rng.integers(0, N, size=6)]))
    rows, sibs = b.open_rows(idx)
    for r, s, i in zip(rows, sibs, idx):
        assert oracle.merkle_verify(r, i, cap, s), i
    wN = oracle.primitive_root_of_unity(n_log + rate_bits)
    for L in (idx[0], idx[4], idx[-1]):
        x = 7 * oracle.exp(wN, oracle.reverse_bits(L, n_log + rate_bits)) % P
        for c in (0, Pn - 1):
            assert int(rows[idx.index(L)][c]) == int(oracle.naive_coset_eval(coeffs_of(c), 0, int(x))[0]), (L, c)
    return cap, rows


@pytest.mark.parametrize("n_log,Pn", [(22, 135), (21, 234), (20, 400)])
def test_scale_sweep_shapes_single_gpu(ctx, n_log, Pn):
    n = 1 << n_log
    vals = p2b.DeviceBuffer(ctx, Pn * n)
    ctx.fill_synthetic(vals, Pn * n, SEED)
    holder = {}

    def factory(n_log, Pn):
        b = p2b.PolynomialBatch.from_values(ctx, (vals, Pn, n), 3, 4)
        holder["b"] = b
        ptrs = b.device_ptrs()

        def coeffs_of(c):
            out = np.empty(n, dtype=np.uint64)
            p2b._check(p2b.lib().p2b_memcpy_d2h(ctx.handle, out.ctypes.data, C.c_void_p(ptrs["coeffs"] + 8 * c * n), 8 * n))
            return out
        return b, coeffs_of
    cap, _ = _commit_properties(ctx, n_log, Pn, factory)
    b = holder["b"]
    # idempotence: committing the device-resident coefficients gives the same cap
    ptrs = b.device_ptrs()

    class _Ptr:
        def __init__(self, ptr):
            self.ptr = ptr
    b2 = p2b.PolynomialBatch.from_coeffs(ctx, (_Ptr(ptrs["coeffs"]), Pn, n), 3, 4)
    assert np.array_equal(b2.cap(), cap)
    b2.close()
    b.close()
    vals.free()


@pytest.mark.parametrize("n_log,Pn", [(22, 135), (20, 400)])
def test_scale_sweep_shapes_multi_device_one_process(ctx, n_log, Pn):
    import torch
    G = 1
    while G * 2 <= min(torch.cuda.device_count(), 8):
        G *= 2
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    from plonky2_gpu_b200 import sharded
    n = 1 << n_log
    mg = p2b.MultiGpu(count=G)
    L = p2b.lib()
    for d in range(G):
        ptr, rounds = mg.resident_cols(d, n_log, Pn)
        ctxh = L.p2b_mgpu_ctx(mg.handle, d)
        for row0, c0, c1 in sharded.local_layout(Pn, G, d)[1]:
            if c1 > c0:
                p2b._check(L.p2b_fill_synthetic(ctxh, ptr + row0 * n * 8, (c1 - c0) * n, SEED, c0 * n))
    mb = mg.commit_resident(n_log, Pn, 3, 4)
    cap_m = mb.cap()
    vals = p2b.DeviceBuffer(ctx, Pn * n)
    ctx.fill_synthetic(vals, Pn * n, SEED)
    b = p2b.PolynomialBatch.from_values(ctx, (vals, Pn, n), 3, 4)
    assert np.array_equal(b.cap(), cap_m)
    N = n << 3
    idx = [0, N - 1, N // G, N // G - 1, 3 * (N // G) + 5 if G > 3 else 5]
    r1, s1 = b.open_rows(idx)
    r2, s2 = mb.open_rows(idx)
    assert np.array_equal(r1, r2) and np.array_equal(s1, s2)
    b.close()
    mb.close()
    mg.close()
    vals.free()


def test_widest_shape_the_reference_cuda_abi_can_address(ctx):
    """2^20 x 234 (N * P = 1.96e9 < 2^31, the reference passes sizes as C int): cap == the reference's own kernels."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_cuda_bench
    if not os.path.exists(ref_cuda_bench.REF_SO):
        pytest.skip("oracle/_ref/libplonky2_ref_cuda.so not built")
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)       # the reference printf()s its kernel timings
    try:
        r = ref_cuda_bench.measure(20, 234, reps=1, ctx=ctx, with_ours=True)
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    assert r["caps_equal"], r
