"""GPU: BASELINE.json full-size workloads checked through size-independent properties (the oracle would need minutes):
  * opened rows verify against the cap with the CPU verifier (verify_merkle_proof_to_cap, merkle_proofs.rs:53-81),
  * an opened row equals the direct evaluation of the returned coefficients at that LDE point (oracle Horner),
  * from_coeffs(from_values(v).coefficients) reproduces the same cap (idempotence of the commit on its own output),
  * the inverse NTT round-trips: evaluating the coefficients on H (rows of a rate-0 commit) returns the input values.
"""
import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
P = oracle.ORDER
SEED = 0x504C4F4E4B5932


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


def synth(ctx, Pn, n):
    d = p2b.DeviceBuffer(ctx, Pn * n)
    ctx.fill_synthetic(d, Pn * n, SEED)
    ctx.synchronize()
    return d


def eval_at(coeffs_col, x):
    # Horner in Python ints on a strided subsample is too slow for 2^20; use the oracle's naive evaluator on one point:
    # naive_coset_eval(c, n_log=0, shift=x) evaluates at x * <w_1> = {x}
    return int(oracle.naive_coset_eval(coeffs_col, 0, int(x))[0])


@pytest.mark.parametrize("n_log,Pn", [(20, 135), (22, 20)])
def test_full_size_commit_properties(ctx, n_log, Pn):
    rate_bits, cap_height = 3, 4
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    vals = synth(ctx, Pn, n)
    b = p2b.PolynomialBatch.from_values(ctx, (vals, Pn, n), rate_bits, cap_height)
    cap = b.cap()
    rng = np.random.default_rng(5)
    idx = sorted(set([0, 1, N - 1, n - 1, n, 3 * n + 17] + [int(x) for x in rng.integers(0, N, size=10)]))
    rows, sibs = b.open_rows(idx)
    for r, s, i in zip(rows, sibs, idx):
        assert oracle.merkle_verify(r, i, cap, s), i
    # rows are evaluations of the coefficient polynomials: leaf L <-> point g * w_N^(reverse_bits(L))
    coeffs = b.polynomials()
    wN = oracle.primitive_root_of_unity(n_log + rate_bits)
    for L in (idx[0], idx[3], idx[-1]):
        x = 7 * oracle.exp(wN, oracle.reverse_bits(L, n_log + rate_bits)) % P
        for c in (0, Pn // 2, Pn - 1):
            assert int(rows[idx.index(L)][c]) == eval_at(coeffs[c], x), (L, c)
    # idempotence: committing the returned coefficients gives the same cap
    b2 = p2b.PolynomialBatch.from_coeffs(ctx, coeffs, rate_bits, cap_height)
    assert np.array_equal(b2.cap(), cap)
    b2.close()
    b.close()
    # inverse NTT round trip on two columns: coefficient evaluation on the subgroup returns the synthetic values
    v_host = vals.to_host(2 * n).reshape(2, n)
    wn = oracle.primitive_root_of_unity(n_log)
    for c in range(2):
        for j in (0, 1, n // 2 + 3, n - 1):
            assert int(v_host[c][j]) == eval_at(coeffs[c], oracle.exp(wn, j))


def test_wide_ecc_shape_and_quotient_scale(ctx):
    """BASELINE config 4 shape: ~2^17 rows, 234 wires (wide_ecc_config): commit + proof verification."""
    n_log, Pn, rate_bits, cap_height = 17, 234, 3, 4
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    vals = synth(ctx, Pn, n)
    b = p2b.PolynomialBatch.from_values(ctx, (vals, Pn, n), rate_bits, cap_height)
    cap = b.cap()
    idx = [0, 12345, N - 1]
    rows, sibs = b.open_rows(idx)
    for r, s, i in zip(rows, sibs, idx):
        assert oracle.merkle_verify(r, i, cap, s)
    b.close()


def test_full_size_fri_proof_is_accepted_by_the_verifier(ctx):
    """BASELINE-scale FRI (2^18 x 259 polynomials shaped like the four oracles, standard config: rate 3, cap 4, 16 PoW
    bits, 28 queries, arity 16): the restated verifier (fri/verifier.rs) accepts the device's proof; a tampered one fails."""
    from oracle import fri as FR
    n_log, rate_bits, cap_height, pow_bits, queries = 18, 3, 4, 16, 28
    polys = (88, 135, 20, 16)
    n = 1 << n_log
    arity = FR.constant_arity_bits(4, 5, n_log, rate_bits, cap_height)
    batches_dev = []
    for k in polys:
        d = synth(ctx, k, n)
        batches_dev.append(p2b.PolynomialBatch.from_values(ctx, (d, k, n), rate_bits, cap_height))
        del d
    zeta = (0x123456789abcdef, 0xfedcba987654321)
    g = oracle.primitive_root_of_unity(n_log)
    zeta_next = FR.escale(zeta, g)
    all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
    batches = [FR.FriBatchInfo(zeta, all_polys), FR.FriBatchInfo(zeta_next, [(2, 0), (2, 1)])]
    params = FR.FriParams(n_log, rate_bits, cap_height, pow_bits, queries, arity)
    start = FR.Challenger([3] * 12, [1, 2, 3])
    gch = p2b.Challenger(start.sponge_state, start.input_buffer, start.output_buffer)
    proof = p2b.fri_prove_openings(ctx, batches_dev, [(b.point, b.polynomials) for b in batches], gch, n_log, rate_bits, cap_height,
                                   pow_bits, queries, arity)
    openings = []
    for b in batches:
        per = [p2b.eval_openings(ctx, gb, b.point) for gb in batches_dev]
        openings.append([tuple(int(x) for x in per[o][p]) for (o, p) in b.polynomials])
    pr = FR.FriProof()
    pr.commit_phase_merkle_caps = proof.commit_phase_merkle_caps
    pr.final_poly = [(int(a), int(c)) for a, c in proof.final_poly]
    pr.pow_witness = proof.pow_witness
    pr.query_round_proofs = [([(rows[q], sibs[q]) for rows, sibs in proof.initial], [(ev[q].reshape(-1), sibs[q]) for ev, sibs in proof.steps])
                             for q in range(queries)]
    challenges = FR.fri_challenges(start.clone(), pr.commit_phase_merkle_caps, pr.final_poly, pr.pow_witness, params)
    assert 64 - challenges[2].bit_length() >= pow_bits
    caps = [gb.cap() for gb in batches_dev]
    assert FR.verify_fri_proof(batches, (False,) * 4, openings, challenges, caps, pr, params)
    assert len(pr.final_poly) == 1 << (n_log - sum(arity))
    # tampering with one opened value breaks the verification
    bad = list(openings[0])
    bad[7] = (bad[7][0] ^ 1, bad[7][1])
    with pytest.raises(AssertionError):
        FR.verify_fri_proof(batches, (False,) * 4, [bad, openings[1]], challenges, caps, pr, params)
    proof.close()
    for gb in batches_dev:
        gb.close()
