"""GPU: cross-check against the REFERENCE ITSELF.  oracle/_ref/libplonky2_ref_cuda.so is the reference's own CUDA
translation unit (cuda/plonky2_gpu.cu) compiled unmodified for sm_100a by oracle/Makefile (`make ref`, in the build
container where /root/reference exists; the .so travels to the GPU box).  Its `ifft` and `merkle_tree_from_coeffs` are
run on the same inputs as (i) the CPU oracle and (ii) this library's drop-in symbols of the same name, with the
reference's device-memory layout (cuda/plonky2_gpu.cu:435-606, fri/oracle.rs:316-455).  This pins the oracle's LDE
values, digests and caps to outputs of the reference, not only to its Poseidon/bit-reversal known answers."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libplonky2_ref_cuda.so")
P_ = oracle.ORDER


class RefStreams(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("stream2", C.c_void_p)]


def canon(a):
    a = a.copy()
    a[a >= np.uint64(P_)] -= np.uint64(P_)
    return a


@pytest.fixture(scope="module")
def env():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libplonky2_ref_cuda.so not built (needs /root/reference at build time)")
    import torch
    p2b.build()
    ctx = p2b.Context(0)
    ref = C.CDLL(REF_SO)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    streams = RefStreams(s1.cuda_stream, s2.cuda_stream)
    yield ctx, ref, streams, (s1, s2)
    ctx.close()


def run_layout(fn_ifft, fn_tree, ctx, streams, values, rate_bits, cap_height):
    """Drive `ifft` + `merkle_tree_from_coeffs` the way fri/oracle.rs:352-457 does; returns coeffs, leaves, digests, cap."""
    Pn, n = values.shape
    n_log = n.bit_length() - 1
    N = n << rate_bits
    ncap = 1 << cap_height
    nd = 2 * (N - ncap)
    pad = N * Pn
    total = 2 * pad + 4 * (nd + ncap)
    base = p2b.DeviceBuffer(ctx, total)
    p2b._check(p2b.lib().p2b_memcpy_h2d(ctx.handle, base.ptr, values.ctypes.data, values.size * 8))
    root1 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log))
    root2 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log + rate_bits))
    sp = np.empty(n, dtype=np.uint64)
    cur = 1
    for i in range(n):
        sp[i] = cur
        cur = cur * 7 % P_
    shift = p2b.DeviceBuffer.from_host(ctx, sp)
    n_inv = C.c_uint64(oracle.inverse_2exp(n_log))
    fn_ifft.restype = p2b.RustError
    fn_tree.restype = p2b.RustError
    e = fn_ifft(C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), C.c_void_p(root1.ptr), C.byref(n_inv), C.byref(streams))
    assert e.code == 0
    coeffs = base.to_host(Pn * n).reshape(Pn, n)
    e = fn_tree(C.c_void_p(base.ptr), C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), C.c_void_p(root1.ptr),
                C.c_void_p(root2.ptr), C.c_void_p(shift.ptr), C.c_int(rate_bits), C.c_int(0), C.c_int(cap_height), C.c_int(pad),
                C.byref(streams))
    assert e.code == 0
    ctx.synchronize()
    leaves = base.to_host(N * Pn).reshape(N, Pn)
    dig = base.to_host(4 * nd, offset=2 * pad).reshape(nd, 4)
    cap = base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4)
    return canon(coeffs), canon(leaves), canon(dig), canon(cap)


@pytest.mark.parametrize("n_log,Pn,cap_height", [(9, 5, 4), (10, 7, 4), (12, 20, 4), (11, 9, 0), (13, 135, 4)])
def test_reference_cuda_vs_oracle_vs_ours(env, n_log, Pn, cap_height):
    ctx, ref, streams, _ = env
    rng = np.random.default_rng(4242 + n_log)
    values = rng.integers(0, P_, size=(Pn, 1 << n_log), dtype=np.uint64)
    # Pn >= 5: the reference CUDA hashes leaves of <= 4 elements too, unlike its CPU path (hash_or_noop, config.rs:56-67)
    rate_bits = 3  # the reference kernels hard-code 2^3 - 1 zero blocks (plonky2_gpu_impl.cuh:290-294)
    r_coeffs, r_leaves, r_dig, r_cap = run_layout(ref.ifft, ref.merkle_tree_from_coeffs, ctx, streams, values, rate_bits, cap_height)
    L = p2b.lib()
    o_coeffs, o_leaves, o_dig, o_cap = run_layout(L.ifft, L.merkle_tree_from_coeffs, ctx, streams, values, rate_bits, cap_height)
    e = oracle.batch_from_values(values, rate_bits, cap_height)
    # reference GPU == oracle (pins the oracle to the reference's own outputs)
    assert np.array_equal(r_coeffs, e.coeffs)
    assert np.array_equal(r_leaves, e.leaves)
    assert np.array_equal(r_cap, e.cap)
    assert np.array_equal(r_dig, e.digests)
    # this library's drop-in symbols == reference GPU, in the reference's memory layout
    assert np.array_equal(o_coeffs, r_coeffs)
    assert np.array_equal(o_leaves, r_leaves)
    assert np.array_equal(o_dig, r_dig)
    assert np.array_equal(o_cap, r_cap)


def test_compat_build_merkle_tree_and_transpose(env):
    """build_merkle_tree (lib.rs:71-81) + transpose (plonky2_gpu.cu:192-215) on a column-major LDE in the work area."""
    ctx, ref, streams, _ = env
    rng = np.random.default_rng(7)
    n_log, Pn, rate_bits, cap_height = 8, 6, 2, 3
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    coeffs = rng.integers(0, P_, size=(Pn, n), dtype=np.uint64)
    lde_cols = np.stack([oracle.lde_coset_fft(coeffs[c], rate_bits) for c in range(Pn)])  # natural order, column-major
    pad = N * Pn
    ncap, nd = 1 << cap_height, 2 * (N - (1 << cap_height))
    base = p2b.DeviceBuffer(ctx, 2 * pad + 4 * (nd + ncap))
    p2b._check(p2b.lib().p2b_memcpy_h2d(ctx.handle, base.ptr + 8 * pad, lde_cols.ctypes.data, lde_cols.size * 8))
    L = p2b.lib()
    e = L.build_merkle_tree(base.ptr, Pn, n, n_log, rate_bits, 0, cap_height, pad, C.byref(streams))
    assert e.code == 0
    e = L.transpose(base.ptr, Pn, n, rate_bits, 0, pad, C.byref(streams))
    assert e.code == 0
    exp = oracle.batch_from_coeffs(coeffs, rate_bits, cap_height)
    assert np.array_equal(base.to_host(N * Pn).reshape(N, Pn), exp.leaves)
    assert np.array_equal(base.to_host(4 * nd, offset=2 * pad).reshape(nd, 4), exp.digests)
    assert np.array_equal(base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4), exp.cap)


# ---- gate constraints: the reference's own device-side evaluators (cuda/*Gate.cuh) vs the oracle's restatement -------
def _gate_cases():
    from oracle import quotient as Q
    return [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate(), Q.ArithmeticGate(20), Q.BaseSumGate(63, 2), Q.BaseSumGate(16, 4),
            Q.PoseidonGate(), Q.RandomAccessGate(4, 4, 2), Q.RandomAccessGate(2, 13, 2), Q.U32ArithmeticGate(6), Q.U32AddManyGate(3, 9),
            Q.U32AddManyGate(11, 5), Q.U32RangeCheckGate(8), Q.U32RangeCheckGate(1), Q.U32SubtractionGate(11), Q.ComparisonGate(32, 16),
            Q.ComparisonGate(8, 4)]


@pytest.mark.parametrize("gate", _gate_cases(), ids=lambda g: type(g).__name__ + str(g.params))
def test_gate_constraints_match_reference_cuda_gates(env, gate):
    """Pins oracle/quotient.py's gate restatements to the reference's code: same rows through cuda/<Gate>.cuh's
    eval_unfiltered_base_packed (instantiated like cuda/plonky2_gpu_impl.cuh:633-685) and through the oracle."""
    from tests import quotient_fixtures as F
    ctx, ref, streams, _ = env
    ref.p2ref_eval_gate.restype = C.c_int
    ref.p2ref_eval_gate.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    rng = np.random.default_rng(5)
    nw, ncst, rows = 234, 4, 24
    pih = [int(x) for x in rng.integers(0, P_, size=4, dtype=np.uint64)]
    wires = np.zeros((rows, nw), dtype=np.uint64)
    consts = np.zeros((rows, ncst), dtype=np.uint64)
    for r in range(rows):
        gc = [int(x) for x in rng.integers(0, P_, size=ncst, dtype=np.uint64)]
        if r % 2 == 0:      # a row the gate's generator would write (most constraints vanish) ...
            wires[r] = np.array(F.honest_row(gate, rng, nw, gc, pih), dtype=np.uint64)
        else:               # ... and arbitrary field elements (every constraint exercised with non-trivial values)
            wires[r] = rng.integers(0, P_, size=nw, dtype=np.uint64)
        consts[r] = np.array(gc, dtype=np.uint64)
    nout = max(gate.num_constraints(), 1)
    dw, dk = p2b.DeviceBuffer.from_host(ctx, wires.reshape(-1)), p2b.DeviceBuffer.from_host(ctx, consts.reshape(-1))
    dp = p2b.DeviceBuffer.from_host(ctx, np.array(pih, dtype=np.uint64))
    do = p2b.DeviceBuffer(ctx, rows * nout)
    ctx.synchronize()
    params = list(gate.params) + [0, 0, 0]
    rc = ref.p2ref_eval_gate(gate.type_id, params[0], params[1], params[2], dw.ptr, nw, dk.ptr, ncst, dp.ptr, do.ptr, nout, rows)
    assert rc == 0
    got = canon(do.to_host().reshape(rows, nout))[:, :gate.num_constraints()]
    for r in range(rows):
        want = gate.eval_unfiltered([int(x) for x in consts[r]], [int(x) for x in wires[r]], pih)
        assert [int(x) for x in got[r]] == want, "row %d" % r


# ---- quotient values: the reference's own kernel (hard-wired to its 25-gate circuit) vs the oracle vs this library ----
def _reference_circuit(degree_bits):
    """The circuit cuda/plonky2_gpu_impl.cuh:596-685 is compiled for: 25 gate instances, 6 selector groups."""
    from oracle import quotient as Q
    gates = [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate(), Q.BaseSumGate(32, 2), Q.BaseSumGate(63, 2), Q.ArithmeticGate(20),
             Q.BaseSumGate(16, 4), Q.ComparisonGate(32, 16), Q.U32AddManyGate(0, 11), Q.U32AddManyGate(11, 5), Q.U32AddManyGate(13, 5),
             Q.U32AddManyGate(15, 4), Q.U32AddManyGate(16, 4), Q.U32AddManyGate(2, 10), Q.U32AddManyGate(3, 9), Q.U32AddManyGate(5, 9),
             Q.U32AddManyGate(7, 8), Q.U32AddManyGate(9, 6), Q.U32ArithmeticGate(6), Q.U32RangeCheckGate(0), Q.U32RangeCheckGate(1),
             Q.U32RangeCheckGate(8), Q.U32SubtractionGate(11), Q.RandomAccessGate(4, 4, 2), Q.PoseidonGate()]
    sel = [0] * 6 + [1] * 5 + [2] * 5 + [3] * 5 + [4] * 3 + [5]
    groups = [(0, 6), (6, 11), (11, 16), (16, 21), (21, 24), (24, 25)]
    k_is = [pow(7, j, P_) for j in range(80)]
    circ = Q.Circuit(gates, sel, groups, 234, 80, 8, k_is, degree_bits, 3, 2, 8)
    assert circ.num_gate_constraints == 231          # the constant the reference kernel asserts (plonky2_gpu_impl.cuh:512-514)
    return circ


def test_quotient_values_match_reference_cuda_kernel(env):
    """compute_quotient_values_kernel (cuda/plonky2_gpu_impl.cuh:485-876) on random leaves of its own circuit: the oracle's
    quotient values and this library's quotient kernel reproduce it bit-for-bit -- every gate filter, the permutation
    argument, L_0, the alpha-reduction order and the division by Z_H are pinned to the reference's code."""
    from oracle import quotient as Q
    ctx, ref, streams, _ = env
    degree_bits = 4
    circ = _reference_circuit(degree_bits)
    N = 1 << (degree_bits + 3)
    rng = np.random.default_rng(77)
    rnd = lambda shape: rng.integers(0, P_, size=shape, dtype=np.uint64)
    wires, zs_pp, cs = rnd((N, 234)), rnd((N, 2 * (1 + circ.num_partial_products))), rnd((N, 8 + 80))
    cs[:, :6] = rng.integers(0, 25, size=(N, 6), dtype=np.uint64)     # selector columns: small values like real ones
    pih = [int(x) for x in rnd(4)]
    alphas, betas, gammas = ([int(x) for x in rnd(2)] for _ in range(3))
    # the tables the Rust side uploads (fri/oracle.rs + plonk/prover.rs:535-568): points w^i, Z_H on the coset and inverses
    w = oracle.primitive_root_of_unity(degree_bits + 3)
    points = np.array([pow(w, i, P_) for i in range(N)], dtype=np.uint64)
    g_pow_n = pow(7, 1 << degree_bits, P_)
    v = oracle.primitive_root_of_unity(3)
    zh = [(g_pow_n * pow(v, i, P_) - 1) % P_ for i in range(8)]
    zh_inv = [Q.inv(z) for z in zh]
    dev = lambda a: p2b.DeviceBuffer.from_host(ctx, np.ascontiguousarray(a, dtype=np.uint64).reshape(-1))
    d_w, d_z, d_c, d_pts = dev(wires), dev(zs_pp), dev(cs), dev(points)
    d_zh, d_zhi, d_k = dev(np.array(zh, dtype=np.uint64)), dev(np.array(zh_inv, dtype=np.uint64)), dev(np.array(circ.k_is, dtype=np.uint64))
    d_a, d_b, d_g = (dev(np.array(x, dtype=np.uint64)) for x in (alphas, betas, gammas))
    d_out = p2b.DeviceBuffer(ctx, 2 * N)
    ctx.synchronize()
    ref.p2ref_quotient_values.restype = C.c_int
    ref.p2ref_quotient_values.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    rc = ref.p2ref_quotient_values(degree_bits, d_pts.ptr, d_out.ptr, (C.c_uint64 * 4)(*pih), d_c.ptr, 88, d_z.ptr, zs_pp.shape[1], d_w.ptr, 234,
                                   8, circ.num_partial_products, d_zh.ptr, d_zhi.ptr, d_k.ptr, d_a.ptr, d_b.ptr, d_g.ptr)
    assert rc == 0
    ref_vals = canon(d_out.to_host().reshape(N, 2))
    # oracle
    want = Q.compute_quotient_values(circ, wires, zs_pp, cs, pih, betas, gammas, alphas)
    assert [[int(x) for x in r] for r in ref_vals] == [list(r) for r in want]
    # this library, same rows
    pc = p2b.Circuit([(g.type_id, g.params) for g in circ.gates], circ.selector_indices, circ.groups, 234, 80, 8, circ.k_is, degree_bits, 3, 2, 8)
    vals, _ = p2b.compute_quotient_polys_rows(ctx, pc, wires, zs_pp, cs, pih, betas, gammas, alphas)
    assert np.array_equal(vals.T, ref_vals)
