"""CPU: the FRI restatement (oracle/fri.py) is self-consistent the way the reference's own tests check the path
(prove -> verify, plonky2/src/fri/mod.rs + plonk/proof.rs tests): the restated verifier accepts the restated prover's
proof and rejects corrupted ones; extension-field and challenger identities."""
import numpy as np
import pytest

import oracle
from oracle import fri as FR
from oracle.quotient import P
from tests.fri_fixtures import make_instance


def run_verify(oracles, batches, params, ch_verifier, proof, salted):
    openings = FR.fri_openings(batches, oracles)
    challenges = FR.fri_challenges(ch_verifier, proof.commit_phase_merkle_caps, proof.final_poly, proof.pow_witness, params)
    return FR.verify_fri_proof(batches, salted, openings, challenges, [o.cap for o in oracles], proof, params)


def test_extension_field_identities():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))
        b = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))
        assert FR.emul(a, FR.einv(a)) == (1, 0)
        assert FR.emul(a, b) == FR.emul(b, a)
        assert FR.epow(a, 5) == FR.emul(FR.emul(FR.emul(a, a), FR.emul(a, a)), a)
    assert FR.emul((0, 1), (0, 1)) == (7, 0)  # X^2 = W = 7 (goldilocks_extensions.rs:19)
    # DTH_ROOT = W^((p-1)/2) = p - 1 (goldilocks_extensions.rs:22): frobenius negates c1
    assert pow(7, (P - 1) // 2, P) == 18446744069414584320


def test_challenger_matches_hash_no_pad_sponge():
    # absorbing 8k elements then squeezing = the overwrite-mode sponge of hashing.rs:81-104: state[0..4] after the
    # last permutation is hash_n_to_hash_no_pad of the same elements; challenges pop from the END of the rate part
    rng = np.random.default_rng(3)
    xs = [int(x) for x in rng.integers(0, P, size=16, dtype=np.uint64)]
    ch = FR.Challenger()
    ch.observe_elements(xs)
    assert ch.input_buffer == [] and ch.output_buffer[:4] == [int(x) for x in oracle.hash_no_pad(np.array(xs, dtype=np.uint64))]
    first = ch.get_challenge()
    assert first == ch.sponge_state[7]
    # a partially filled buffer forces a duplexing on the next challenge, overwriting only the first lanes
    ch2 = FR.Challenger()
    ch2.observe_elements(xs[:3])
    c = ch2.get_challenge()
    st = np.zeros(12, dtype=np.uint64)
    st[:3] = xs[:3]
    assert c == int(oracle.poseidon(st)[7])


def test_divide_by_linear_and_final_poly_vanish():
    oracles, batches, params, ch = make_instance()
    alpha = (123456789, 987654321)
    fin = FR.final_poly_coeffs(batches, oracles, alpha)
    n = 1 << params.degree_bits
    assert len(fin) == n and fin[0] == (0, 0)
    # final(X)/X * prod (X - z_i) must equal the alpha-combination of (F_i - F_i(z_i)) * (X - z_other): check by
    # evaluating at a random extension point
    x = (5, 11)
    acc = (0, 0)
    for c in reversed(fin):
        acc = FR.eadd(FR.emul(acc, x), c)
    lhs = FR.emul(acc, FR.einv(x))
    terms = []
    for b in batches:
        comp_x, comp_z, ap = (0, 0), (0, 0), (1, 0)
        for (o, p) in b.polynomials:
            comp_x = FR.eadd(comp_x, FR.emul(ap, FR.eval_poly_base_at_ext(oracles[o].coeffs[p], x)))
            comp_z = FR.eadd(comp_z, FR.emul(ap, FR.eval_poly_base_at_ext(oracles[o].coeffs[p], b.point)))
            ap = FR.emul(ap, alpha)
        terms.append(FR.emul(FR.esub(comp_x, comp_z), FR.einv(FR.esub(x, b.point))))
    rhs = FR.eadd(FR.emul(terms[0], FR.epow(alpha, len(batches[1].polynomials))), terms[1])
    assert lhs == rhs


@pytest.mark.parametrize("salted", [(False,) * 4, (False, True, True, True)])
def test_prove_then_verify(salted):
    oracles, batches, params, ch = make_instance(salted=salted)
    ch_v = ch.clone()
    proof = FR.prove_openings(batches, oracles, ch, params)
    assert len(proof.final_poly) == params.final_poly_len
    assert len(proof.commit_phase_merkle_caps) == len(params.reduction_arity_bits)
    assert run_verify(oracles, batches, params, ch_v, proof, salted)
    # prover and verifier transcripts end in the same state
    assert ch.sponge_state == ch_v.sponge_state


def test_verifier_rejects_corruption():
    salted = (False,) * 4
    oracles, batches, params, ch = make_instance(seed=2)
    base = ch.clone()
    proof = FR.prove_openings(batches, oracles, ch, params)
    # corrupt one evaluation of the first FRI step of query 0
    flat, sib = proof.query_round_proofs[0][1][0]
    flat = flat.copy()
    flat[0] ^= np.uint64(1)
    good = proof.query_round_proofs[0]
    proof.query_round_proofs[0] = (good[0], [(flat, sib)] + good[1][1:])
    with pytest.raises(AssertionError):
        run_verify(oracles, batches, params, base.clone(), proof, salted)
    proof.query_round_proofs[0] = good
    # corrupt the final polynomial
    fp = list(proof.final_poly)
    proof.final_poly = [(fp[0][0] ^ 1, fp[0][1])] + fp[1:]
    with pytest.raises(AssertionError):
        run_verify(oracles, batches, params, base.clone(), proof, salted)
    proof.final_poly = fp
    # a wrong proof-of-work witness
    w = proof.pow_witness
    proof.pow_witness = w + 1
    with pytest.raises(AssertionError):
        run_verify(oracles, batches, params, base.clone(), proof, salted)
    proof.pow_witness = w
    assert run_verify(oracles, batches, params, base.clone(), proof, salted)


def test_constant_arity_bits_strategy():
    # fri/reduction_strategies.rs:38-49 on the standard recursion config (arity 4 bits, final poly 5 bits)
    assert FR.constant_arity_bits(4, 5, 20, 3, 4) == [4, 4, 4, 4]
    assert FR.constant_arity_bits(4, 5, 12, 3, 4) == [4, 4]
    assert FR.constant_arity_bits(4, 5, 5, 3, 4) == []
    assert FR.constant_arity_bits(3, 0, 4, 1, 4) == []  # would leave a tree shorter than the cap
