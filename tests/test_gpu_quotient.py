"""GPU: compute_quotient_polys on the device (through the C ABI) against the Python restatement of
plonky2/src/plonk/prover.rs:790-1034 -- bit-exact quotient evaluations and coefficients."""
import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from oracle import quotient as Q
from tests import quotient_fixtures as F

pytestmark = pytest.mark.gpu
P = Q.P


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


def to_p2b_circuit(circ):
    gates = [(g.type_id, g.params) for g in circ.gates]
    return p2b.Circuit(gates, circ.selector_indices, circ.groups, circ.num_wires, circ.num_routed_wires, circ.num_constants,
                       circ.k_is, circ.degree_bits, circ.rate_bits, circ.num_challenges, circ.quotient_degree_factor)


def run_both(ctx, inst, cap_height=0):
    c = inst.circ
    bw = p2b.PolynomialBatch.from_values(ctx, inst.wires, c.rate_bits, cap_height)
    bz = p2b.PolynomialBatch.from_values(ctx, inst.zs_pp, c.rate_bits, cap_height)
    bc = p2b.PolynomialBatch.from_values(ctx, inst.consts_sigmas, c.rate_bits, cap_height)
    vals, coeffs = p2b.compute_quotient_polys(ctx, to_p2b_circuit(c), bw, bz, bc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    ow = oracle.batch_from_values(inst.wires, c.rate_bits, 0, want_digests=False).leaves
    oz = oracle.batch_from_values(inst.zs_pp, c.rate_bits, 0, want_digests=False).leaves
    oc = oracle.batch_from_values(inst.consts_sigmas, c.rate_bits, 0, want_digests=False).leaves
    evals, ecoeffs = Q.compute_quotient_polys(c, ow, oz, oc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    for b in (bw, bz, bc):
        b.close()
    return vals, coeffs, evals, ecoeffs


@pytest.mark.parametrize("which,degree_bits", [(0, 4), (1, 4), (1, 6)])
def test_quotient_honest_witness_matches_oracle(ctx, which, degree_bits):
    gates, groups, sel = F.standard_gate_sets()[which]
    inst = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=21 + which)
    vals, coeffs, evals, ecoeffs = run_both(ctx, inst)
    n = 1 << degree_bits
    for c in range(inst.circ.num_challenges):
        assert np.array_equal(vals[c], evals[c])
        assert np.array_equal(coeffs[c], ecoeffs[c])
        assert not np.any(coeffs[c][7 * n:])  # honest witness: degree < 7n (size-independent property)


@pytest.mark.parametrize("degree_bits,random_data", [(4, False), (5, True)])
def test_quotient_recursion_gate_set_matches_oracle(ctx, degree_bits, random_data):
    # the gates a recursive-verifier circuit adds: extension arithmetic, reducing, exponentiation, Poseidon MDS and both
    # interpolation gates, on honest rows and on random data (every constraint non-zero)
    gates, groups, sel = F.recursion_gate_set()
    inst = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=81)
    if random_data:
        rng = np.random.default_rng(82)
        for m in (inst.wires, inst.zs_pp):
            m[:] = rng.integers(0, P, size=m.shape, dtype=np.uint64)
        inst.consts_sigmas[inst.circ.num_selectors:] = rng.integers(0, P, size=inst.consts_sigmas[inst.circ.num_selectors:].shape, dtype=np.uint64)
    vals, coeffs, evals, ecoeffs = run_both(ctx, inst)
    n = 1 << degree_bits
    for c in range(inst.circ.num_challenges):
        assert np.array_equal(vals[c], evals[c])
        assert np.array_equal(coeffs[c], ecoeffs[c])
        if not random_data:
            assert not np.any(coeffs[c][7 * n:])  # honest rows: the numerator vanishes on H


def test_quotient_random_data_matches_oracle(ctx):
    # no constraint holds on random data: every term of the alpha reduction is exercised with non-zero values,
    # including real partial products / Z columns and a non-identity sigma
    gates, groups, sel = F.standard_gate_sets()[1]
    inst = F.build_instance(gates, groups, sel, 4, 135, 80, seed=31)
    rng = np.random.default_rng(32)
    for m in (inst.wires, inst.zs_pp, inst.consts_sigmas):
        m[:] = rng.integers(0, P, size=m.shape, dtype=np.uint64)
    vals, coeffs, evals, ecoeffs = run_both(ctx, inst)
    for c in range(inst.circ.num_challenges):
        assert np.array_equal(vals[c], evals[c])
        assert np.array_equal(coeffs[c], ecoeffs[c])


def test_quotient_edge_values_take_the_exact_redo_path(ctx):
    # raw LDE rows filled with the operands that make the optimistic reduction flag its rare case (2^48 * 2^48 = 2^96,
    # eps * eps, (p-1)^2 ...): those points are recomputed by the exact path and must still be bit-identical
    gates, groups, sel = F.standard_gate_sets()[1]
    inst = F.build_instance(gates, groups, sel, 3, 135, 80, seed=71)
    c = inst.circ
    N = c.lde_size if hasattr(c, "lde_size") else 1 << (c.degree_bits + c.rate_bits)
    edge = np.array([0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**48, 2**48 - 1, 2**48 + 1, 2**63, P - 2**32,
                     P - 2**48, 2**64 - 2**33, 2**16, 2**47, 2**49], dtype=np.uint64)
    rng = np.random.default_rng(72)
    mats = []
    for m in (inst.wires, inst.zs_pp, inst.consts_sigmas):
        r = rng.integers(0, P, size=(N, m.shape[0]), dtype=np.uint64)
        pick = rng.integers(0, 2, size=r.shape).astype(bool)
        r[pick] = edge[rng.integers(0, edge.size, size=int(pick.sum()))]
        mats.append(r)
    # keep the selector columns meaningful so that every gate's evaluator runs on some rows
    mats[2][:, :c.num_selectors] = rng.integers(0, len(gates), size=(N, c.num_selectors), dtype=np.uint64)
    challenges = [2**48, P - 1], [2**32 - 1, 2**48], [2**48, 2**48 + 1]
    vals, coeffs = p2b.compute_quotient_polys_rows(ctx, to_p2b_circuit(c), *mats, inst.pih, *challenges)
    evals, ecoeffs = Q.compute_quotient_polys(c, *mats, inst.pih, *challenges)
    for ch in range(c.num_challenges):
        assert np.array_equal(vals[ch], evals[ch])
        assert np.array_equal(coeffs[ch], ecoeffs[ch])


@pytest.mark.parametrize("num_wires,num_routed,nch,qdf,rate_bits", [(136, 80, 2, 8, 3), (234, 80, 2, 8, 3), (40, 12, 3, 4, 2), (60, 9, 1, 8, 3)])
def test_quotient_other_configs(ctx, num_wires, num_routed, nch, qdf, rate_bits):
    # standard_ecc_config (136 wires), wide_ecc_config (234 wires) and odd shapes (routed wires not a multiple of the
    # chunk size, 1 and 3 challenges, quotient degree factor 4)
    gates = [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate(), Q.ArithmeticGate(num_routed // 4),
             Q.U32SubtractionGate(Q.U32SubtractionGate.num_ops_for(num_wires, num_routed)), Q.ComparisonGate(6, 3)]
    groups, sel = [(0, 3), (3, 6)], [0, 0, 0, 1, 1, 1]
    inst = F.build_instance(gates, groups, sel, 4, num_wires, num_routed, seed=41, rate_bits=rate_bits, num_challenges=nch,
                            quotient_degree_factor=qdf)
    rng = np.random.default_rng(42)
    inst.zs_pp[:] = rng.integers(0, P, size=inst.zs_pp.shape, dtype=np.uint64)
    vals, coeffs, evals, ecoeffs = run_both(ctx, inst)
    for c in range(nch):
        assert np.array_equal(vals[c], evals[c])
        assert np.array_equal(coeffs[c], ecoeffs[c])


def test_quotient_error_paths(ctx):
    gates, groups, sel = F.standard_gate_sets()[0]
    inst = F.build_instance(gates, groups, sel, 3, 135, 80, seed=51)
    c = inst.circ
    bw = p2b.PolynomialBatch.from_values(ctx, inst.wires, 3, 0)
    bz = p2b.PolynomialBatch.from_values(ctx, inst.zs_pp, 3, 0)
    bc = p2b.PolynomialBatch.from_values(ctx, inst.consts_sigmas, 3, 0)
    bad = p2b.Circuit([(99, ())], [0], [(0, 1)], 135, 80, c.num_constants, c.k_is, 3)
    with pytest.raises(p2b.P2BError, match="unknown gate type"):
        p2b.compute_quotient_polys(ctx, bad, bw, bz, bc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    wrong_deg = p2b.Circuit([(0, ())], [0], [(0, 1)], 135, 80, c.num_constants, c.k_is, 5)
    with pytest.raises(p2b.P2BError, match="batch shape"):
        p2b.compute_quotient_polys(ctx, wrong_deg, bw, bz, bc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    # quotient degree factor above the rate is refused like prover.rs:809-813
    high = p2b.Circuit([(0, ())], [0], [(0, 1)], 135, 80, c.num_constants, c.k_is, 3, rate_bits=3, quotient_degree_factor=16)
    with pytest.raises(p2b.P2BError, match="higher than the rate"):
        p2b.compute_quotient_polys(ctx, high, bw, bz, bc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    for b in (bw, bz, bc):
        b.close()


def test_legacy_compute_quotient_polys_symbol(ctx):
    """The reference's own symbol (cuda/src/lib.rs:117-143) with device slices, on a registered circuit."""
    import ctypes as C
    gates, groups, sel = F.standard_gate_sets()[1]
    inst = F.build_instance(gates, groups, sel, 4, 135, 80, seed=61)
    c = inst.circ
    rng = np.random.default_rng(62)
    inst.zs_pp[:] = rng.integers(0, P, size=inst.zs_pp.shape, dtype=np.uint64)
    bw = p2b.PolynomialBatch.from_values(ctx, inst.wires, 3, 4)
    bz = p2b.PolynomialBatch.from_values(ctx, inst.zs_pp, 3, 4)
    bc = p2b.PolynomialBatch.from_values(ctx, inst.consts_sigmas, 3, 4)
    pc = to_p2b_circuit(c)
    L = p2b.lib()
    p2b._check(L.p2b_compat_set_circuit(C.byref(pc.struct), (C.c_uint64 * 4)(*inst.pih)))
    N = 1 << (4 + 3)
    dev = lambda xs: p2b.DeviceBuffer.from_host(ctx, np.array(xs, dtype=np.uint64))
    da, db, dg, dk = dev(inst.alphas), dev(inst.betas), dev(inst.gammas), dev(c.k_is)
    d_outs, d_q = p2b.DeviceBuffer(ctx, 2 * N), p2b.DeviceBuffer(ctx, 2 * N)
    sl = lambda ptr, ln: C.byref(p2b.DataSlice(ptr, ln))
    import torch
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    class RefStreams(C.Structure):
        _fields_ = [("stream", C.c_void_p), ("stream2", C.c_void_p)]
    streams = RefStreams(s1.cuda_stream, s2.cuda_stream)
    pw, pz, pcs = bw.device_ptrs()["leaves"], bz.device_ptrs()["leaves"], bc.device_ptrs()["leaves"]
    ctx.synchronize()
    e = L.compute_quotient_polys(pw, 135, 16, 4, None, None, 3, 0, sl(pz, N * bz.leaf_len), sl(pcs, N * bc.leaf_len), d_outs.ptr, d_q.ptr,
                                 sl(None, N), sl(None, 8), sl(None, 8), sl(dk.ptr, 80), sl(da.ptr, 2), sl(db.ptr, 2), sl(dg.ptr, 2),
                                 C.byref(streams))
    assert e.code == 0
    ow = oracle.batch_from_values(inst.wires, 3, 0, want_digests=False).leaves
    oz = oracle.batch_from_values(inst.zs_pp, 3, 0, want_digests=False).leaves
    oc = oracle.batch_from_values(inst.consts_sigmas, 3, 0, want_digests=False).leaves
    evals, ecoeffs = Q.compute_quotient_polys(c, ow, oz, oc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    outs = d_outs.to_host().reshape(N, 2)
    polys = d_q.to_host().reshape(2, N)
    for ch in range(2):
        assert np.array_equal(outs[:, ch], evals[ch])
        assert np.array_equal(polys[ch], ecoeffs[ch])
    for b in (bw, bz, bc):
        b.close()
