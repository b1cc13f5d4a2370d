"""GPU: the device reproduces the committed golden anchors (tests/golden/oracle_golden.json): caps and SHA-256 of the
coefficient / leaf / digest arrays of fixed-seed commits, quotient values, a FRI transcript, and the cap of the headline
2^20 x 135 synthetic matrix that bench.py commits."""
import json
import os

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))
P = oracle.ORDER


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


@pytest.mark.parametrize("want", GOLDEN["commits"], ids=lambda w: "x".join(str(x) for x in w["case"]))
def test_commit_anchor(ctx, want):
    n_log, polys, rate_bits, cap_height, seed, blinding = want["case"]
    rng = np.random.default_rng(seed)
    values = rng.integers(0, P, size=(polys, 1 << n_log), dtype=np.uint64)
    salt = rng.integers(0, P, size=(4, 1 << (n_log + rate_bits)), dtype=np.uint64) if blinding else None
    b = p2b.PolynomialBatch.from_values(ctx, values, rate_bits, cap_height, blinding=blinding, salt=salt)
    assert [[int(x) for x in h] for h in b.cap()] == want["cap"]
    assert MG.sha(b.polynomials()) == want["coeffs_sha256"]
    assert MG.sha(b.leaves()) == want["leaves_sha256"]
    assert MG.sha(b.digests()) == want["digests_sha256"]
    b.close()


@pytest.mark.parametrize("want", GOLDEN["quotients"], ids=lambda w: "-".join(str(x) for x in w["case"]))
def test_quotient_anchor(ctx, want):
    from tests import quotient_fixtures as F
    from tests.test_gpu_quotient import to_p2b_circuit
    which, degree_bits, seed = want["case"]
    sets = F.standard_gate_sets() + (F.recursion_gate_set(),)
    gates, groups, sel = sets[which]
    inst = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=seed)
    c = inst.circ
    bw, bz, bc = (p2b.PolynomialBatch.from_values(ctx, m, c.rate_bits, 0) for m in (inst.wires, inst.zs_pp, inst.consts_sigmas))
    vals, coeffs = p2b.compute_quotient_polys(ctx, to_p2b_circuit(c), bw, bz, bc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    assert MG.sha(vals) == want["values_sha256"] and MG.sha(coeffs) == want["coeffs_sha256"]
    for b in (bw, bz, bc):
        b.close()


@pytest.mark.parametrize("want", GOLDEN["fri"], ids=["default", "standard-shape"])
def test_fri_anchor(ctx, want):
    from tests.fri_fixtures import make_instance
    oracles, batches, params, ch = make_instance(**want["case"])
    g = [p2b.PolynomialBatch.from_values(ctx, o.values, params.rate_bits, params.cap_height, salt=o.salt) for o in oracles]
    gch = p2b.Challenger(ch.sponge_state, ch.input_buffer, ch.output_buffer)
    pr = p2b.fri_prove_openings(ctx, g, [(b.point, b.polynomials) for b in batches], gch, params.degree_bits, params.rate_bits,
                                params.cap_height, params.proof_of_work_bits, params.num_query_rounds, params.reduction_arity_bits)
    assert list(pr.alpha) == want["alpha"] and [list(b) for b in pr.betas] == want["betas"]
    assert pr.pow_witness == want["pow_witness"] and pr.query_indices == want["query_indices"]
    assert [[int(a), int(b)] for a, b in pr.final_poly] == want["final_poly"]
    assert [MG.sha(c) for c in pr.commit_phase_merkle_caps] == want["caps_sha256"]
    assert gch.sponge_state == want["challenger_state_after"]
    pr.close()
    for b in g:
        b.close()


def test_headline_cap_anchor(ctx):
    want = GOLDEN.get("headline")
    if not want:
        pytest.skip("headline cap not generated (tests/golden/make_golden.py --headline)")
    n_log, polys = 20, 135
    d = p2b.DeviceBuffer(ctx, polys << n_log)
    ctx.fill_synthetic(d, polys << n_log, 0x504C4F4E4B5932)
    # the device generator equals the host one the oracle consumed
    head = np.empty(64, dtype=np.uint64)
    p2b._check(p2b.lib().p2b_memcpy_d2h(ctx.handle, head.ctypes.data, d.ptr, 64 * 8))
    assert np.array_equal(head, MG.synthetic(64, 0x504C4F4E4B5932))
    b = p2b.PolynomialBatch.from_values(ctx, (d, polys, 1 << n_log), 3, 4)
    assert [[int(x) for x in h] for h in b.cap()] == want["cap"]
    b.close()
