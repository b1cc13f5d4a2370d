"""GPU: the FRI opening proof on the device (p2b_fri_prove_openings / p2b_eval_openings through the C ABI) against the
CPU restatement of plonky2/src/fri/oracle.rs:1046-1110 + fri/prover.rs:23-260 -- every transcript value, cap, row and
Merkle path bit-exact, and the restated verifier (fri/verifier.rs) accepts the device's proof."""
import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from oracle import fri as FR
from oracle.quotient import P
from tests.fri_fixtures import make_instance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


def commit_all(ctx, oracles, params):
    return [p2b.PolynomialBatch.from_values(ctx, o.values, params.rate_bits, params.cap_height, salt=o.salt) for o in oracles]


def gpu_prove(ctx, gbatches, batches, ch, params):
    gch = p2b.Challenger(ch.sponge_state, ch.input_buffer, ch.output_buffer)
    proof = p2b.fri_prove_openings(ctx, gbatches, [(b.point, b.polynomials) for b in batches], gch, params.degree_bits,
                                   params.rate_bits, params.cap_height, params.proof_of_work_bits, params.num_query_rounds,
                                   params.reduction_arity_bits)
    return proof, gch


def as_oracle_proof(gp):
    """The device proof in the shape oracle.fri.verify_fri_proof reads."""
    pr = FR.FriProof()
    pr.commit_phase_merkle_caps = gp.commit_phase_merkle_caps
    pr.final_poly = [(int(a), int(b)) for a, b in gp.final_poly]
    pr.pow_witness = gp.pow_witness
    pr.query_round_proofs = []
    for q in range(len(gp.query_indices)):
        initial = [(rows[q], sibs[q]) for rows, sibs in gp.initial]
        steps = [(ev[q].reshape(-1), sibs[q]) for ev, sibs in gp.steps]
        pr.query_round_proofs.append((initial, steps))
    return pr


@pytest.mark.parametrize("degree_bits,polys", [(5, (3, 5, 2, 2)), (9, (4, 7, 3, 2)), (0, (1, 1, 1, 1))])
def test_eval_openings_matches_oracle(ctx, degree_bits, polys):
    oracles, batches, params, ch = make_instance(degree_bits=degree_bits, polys=polys, arity_bits=(), cap_height=0)
    g = commit_all(ctx, oracles, params)
    for point in (batches[0].point, batches[1].point, (0, 0), (1, 0), (P - 1, P - 1)):
        for o, gb in zip(oracles, g):
            want = np.array([FR.eval_poly_base_at_ext(c, point) for c in o.coeffs], dtype=np.uint64)
            assert np.array_equal(p2b.eval_openings(ctx, gb, point), want)
    for gb in g:
        gb.close()


@pytest.mark.parametrize("kw", [
    dict(),                                                                           # 2^5, arity 4 then 2
    dict(salted=(False, True, True, True)),                                           # blinded oracles (salt columns in rows)
    dict(degree_bits=8, rate_bits=3, cap_height=4, arity_bits=(4,), pow_bits=10, queries=6, seed=5),   # standard config shape
    dict(degree_bits=7, rate_bits=1, cap_height=0, arity_bits=(1, 2, 3), pow_bits=0, queries=3, seed=6),
    dict(degree_bits=6, rate_bits=2, cap_height=2, arity_bits=(), pow_bits=5, queries=2, seed=7),      # no reduction
    dict(degree_bits=3, rate_bits=3, cap_height=1, arity_bits=(3,), pow_bits=2, queries=2, seed=8),    # folds to a constant
])
def test_prove_openings_matches_oracle_and_verifies(ctx, kw):
    oracles, batches, params, ch = make_instance(**kw)
    salted = kw.get("salted", (False,) * 4)
    g = commit_all(ctx, oracles, params)
    for o, gb in zip(oracles, g):
        assert np.array_equal(gb.cap(), o.cap)
    ch_verifier = ch.clone()
    want = FR.prove_openings(batches, oracles, ch.clone(), params)
    got, gch = gpu_prove(ctx, g, batches, ch, params)
    n = 1 << params.degree_bits
    assert got.alpha == want.alpha
    assert np.array_equal(got.final_poly_in(n), np.array(want.final_poly_in, dtype=np.uint64))
    assert got.betas == want.betas
    for a, b in zip(got.commit_phase_merkle_caps, want.commit_phase_merkle_caps):
        assert np.array_equal(a, b)
    assert np.array_equal(got.final_poly, np.array(want.final_poly, dtype=np.uint64).reshape(-1, 2))
    assert got.pow_witness == want.pow_witness
    assert got.query_indices == want.query_indices
    for q, (initial, steps) in enumerate(want.query_round_proofs):
        for o, (row, sib) in enumerate(initial):
            assert np.array_equal(got.initial[o][0][q], row)
            assert np.array_equal(got.initial[o][1][q], sib)
        for r, (flat, sib) in enumerate(steps):
            assert np.array_equal(got.steps[r][0][q].reshape(-1), flat)
            assert np.array_equal(got.steps[r][1][q], sib)
    # transcript state after the proof = the CPU prover's
    cpu_ch = ch.clone()
    FR.prove_openings(batches, oracles, cpu_ch, params)
    assert gch.sponge_state == cpu_ch.sponge_state
    assert gch.input_buffer == cpu_ch.input_buffer and gch.output_buffer == cpu_ch.output_buffer
    # the verifier accepts the device's proof
    openings = [[tuple(int(x) for x in p2b.eval_openings(ctx, g[o], b.point)[p]) for (o, p) in b.polynomials] for b in batches]
    assert openings == FR.fri_openings(batches, oracles)
    pr = as_oracle_proof(got)
    challenges = FR.fri_challenges(ch_verifier, pr.commit_phase_merkle_caps, pr.final_poly, pr.pow_witness, params)
    assert FR.verify_fri_proof(batches, salted, openings, challenges, [o.cap for o in oracles], pr, params)
    got.close()
    for gb in g:
        gb.close()


def test_larger_instance_verifies(ctx):
    # size-independent property at a size the Python prover would take minutes for: the restated verifier accepts the
    # device's proof (openings from the device too, their parity is covered above)
    oracles, batches, params, ch = make_instance(degree_bits=13, rate_bits=3, cap_height=4, polys=(10, 30, 6, 8), arity_bits=(4, 4),
                                                 pow_bits=14, queries=8, seed=11)
    g = commit_all(ctx, oracles, params)
    ch_verifier = ch.clone()
    got, _ = gpu_prove(ctx, g, batches, ch, params)
    openings = []
    for b in batches:
        per_oracle = [p2b.eval_openings(ctx, gb, b.point) for gb in g]
        openings.append([tuple(int(x) for x in per_oracle[o][p]) for (o, p) in b.polynomials])
    pr = as_oracle_proof(got)
    challenges = FR.fri_challenges(ch_verifier, pr.commit_phase_merkle_caps, pr.final_poly, pr.pow_witness, params)
    assert 64 - challenges[2].bit_length() >= 14
    assert FR.verify_fri_proof(batches, (False,) * 4, openings, challenges, [o.cap for o in oracles], pr, params)
    # a corrupted leaf is rejected
    got.steps[0][0][0, 0, 0] ^= np.uint64(1)
    with pytest.raises(AssertionError):
        FR.verify_fri_proof(batches, (False,) * 4, openings, challenges, [o.cap for o in oracles], as_oracle_proof(got), params)
    got.close()
    for gb in g:
        gb.close()


def test_fri_error_paths(ctx):
    oracles, batches, params, ch = make_instance()
    g = commit_all(ctx, oracles, params)
    gch = p2b.Challenger()
    bl = [(b.point, b.polynomials) for b in batches]
    with pytest.raises(p2b.P2BError, match="out of range"):
        p2b.fri_prove_openings(ctx, g, [((1, 2), [(9, 0)])], gch, 5, 2, 1, 3, 4, [2, 1])
    with pytest.raises(p2b.P2BError, match="do not match"):
        p2b.fri_prove_openings(ctx, g, bl, gch, 6, 2, 1, 3, 4, [2, 1])
    with pytest.raises(p2b.P2BError, match="should be at most"):   # MerkleTree::new's assertion on a too-short FRI tree
        p2b.fri_prove_openings(ctx, g, bl, gch, 5, 2, 6, 3, 4, [2, 1])
    with pytest.raises(p2b.P2BError, match="arity"):
        p2b.fri_prove_openings(ctx, g, bl, gch, 5, 2, 1, 3, 4, [0])
    for gb in g:
        gb.close()
