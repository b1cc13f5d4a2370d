"""CPU: pin the oracle (oracle/p2oracle.c) against every known-answer vector the reference's own tests hold
for this path, and restate the reference's property tests."""
import numpy as np
import pytest

import oracle
from tests.golden.reference_kats import BITREV_256, INVERSE_2EXP, POSEIDON_KATS, POWER_OF_TWO_GENERATOR

P = oracle.ORDER


def rand_field(rng, shape):
    return (rng.integers(0, P, size=shape, dtype=np.uint64, endpoint=False)).astype(np.uint64)


def test_poseidon_known_answers():
    # plonky2/src/hash/poseidon_goldilocks.rs:277-318 (test_vectors)
    for inp, out in POSEIDON_KATS:
        got = oracle.poseidon(np.array(inp, dtype=np.uint64))
        assert [int(x) for x in got] == out


def test_poseidon_fast_equals_naive():
    # check_consistency, plonky2/src/hash/poseidon.rs:733-750
    rng = np.random.default_rng(1)
    for _ in range(50):
        s = rand_field(rng, 12)
        assert np.array_equal(oracle.poseidon(s), oracle.poseidon(s, naive=True))
    s = np.arange(12, dtype=np.uint64)
    assert np.array_equal(oracle.poseidon(s), oracle.poseidon(s, naive=True))


def test_poseidon_noncanonical_inputs():
    # values are any u64 in the reference (to_canonical_u64 only on compare): p + x must hash like x
    s = np.array([3, 0, 7, 0xFFFFFFFE, 1, 2, 3, 4, 5, 6, 7, 8], dtype=np.uint64)
    t = s + np.uint64(P)  # all still < 2^64
    assert np.array_equal(oracle.poseidon(s), oracle.poseidon(t))


def test_bit_reversal_golden_table():
    # plonky2/src/util/mod.rs:78-105
    assert list(oracle.reverse_index_bits(np.array([10, 20, 30, 40], dtype=np.uint64))) == [10, 30, 20, 40]
    assert [int(x) for x in oracle.reverse_index_bits(np.arange(256, dtype=np.uint64))] == BITREV_256
    assert list(oracle.reverse_index_bits(np.array([10], dtype=np.uint64))) == [10]


def test_field_literals():
    for e, v in INVERSE_2EXP.items():
        assert oracle.inverse_2exp(e) == v
        assert oracle.mul(v, 1 << e) == 1
    assert oracle.primitive_root_of_unity(32) == POWER_OF_TWO_GENERATOR
    assert oracle.exp(POWER_OF_TWO_GENERATOR, 1 << 32) == 1
    assert oracle.exp(POWER_OF_TWO_GENERATOR, 1 << 31) == P - 1
    # root(23)^8 == root(20)   
    assert oracle.exp(oracle.primitive_root_of_unity(23), 8) == oracle.primitive_root_of_unity(20)


def test_field_arithmetic_vs_python_bigint():
    # field/src/prime_field_testing.rs style: compare against exact integer arithmetic, incl. edge values
    rng = np.random.default_rng(2)
    edge = [0, 1, 2, P - 1, P - 2, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63]
    vals = edge + [int(x) for x in rng.integers(0, 2**64, size=40, dtype=np.uint64)]
    for a in vals:
        for b in vals:
            assert oracle.add(a, b) == (a + b) % P
            assert oracle.sub(a, b) == (a - b) % P
            assert oracle.mul(a, b) == (a * b) % P
    for a in vals:
        if a % P:
            assert oracle.mul(oracle.inverse(a), a) == 1


def test_fft_equals_naive_and_roundtrip():
    # field/src/fft.rs:243-276 (fft_and_ifft)
    degree, padded = 200, 256
    coeffs = np.array([(i * 1337 % 100) for i in range(degree)] + [0] * (padded - degree), dtype=np.uint64)
    points = oracle.fft(coeffs)
    assert np.array_equal(points, oracle.naive_coset_eval(coeffs, 8, 1))
    assert np.array_equal(oracle.ifft(points), coeffs)
    for r in range(4):
        zt = np.concatenate([coeffs, np.zeros(padded * ((1 << r) - 1), dtype=np.uint64)])
        assert np.array_equal(oracle.fft(zt), oracle.fft(zt, zero_factor=r))


def test_coset_fft_equals_naive_and_roundtrip():
    # field/src/polynomial/mod.rs:482-522
    rng = np.random.default_rng(3)
    k = 8
    c = rand_field(rng, 1 << k)
    shift = 7
    v = oracle.coset_fft(c, shift)
    assert np.array_equal(v, oracle.naive_coset_eval(c, k, shift))
    assert np.array_equal(oracle.coset_ifft(v, shift), c)


def test_lde_is_evaluation_on_big_coset():
    rng = np.random.default_rng(4)
    for k, r in [(3, 1), (5, 3), (6, 2)]:
        c = rand_field(rng, 1 << k)
        assert np.array_equal(oracle.lde_coset_fft(c, r), oracle.naive_coset_eval(c, k + r, 7))


def test_root_table_shape():
    # fft_root_table(n).concat(): row for lg_m=1 has two entries [1, -1]; row k starts at 2^k (k>=1)
    t = oracle.fft_root_table_concat(5)
    assert t.size == 32
    assert int(t[0]) == 1 and int(t[1]) == P - 1
    for k in range(1, 5):
        w = oracle.primitive_root_of_unity(k + 1)
        assert int(t[(1 << k)]) == 1 and int(t[(1 << k) + 1]) == w


def test_sponge_semantics():
    rng = np.random.default_rng(5)
    x = rand_field(rng, 135)
    # manual overwrite-mode sponge with the partial last chunk keeping stale lanes (hashing.rs:88-91)
    st = np.zeros(12, dtype=np.uint64)
    for off in range(0, 135, 8):
        ch = x[off:off + 8]
        st[: ch.size] = ch
        st = oracle.poseidon(st)
    assert np.array_equal(oracle.hash_no_pad(x), st[:4])
    # hash_or_noop (plonk/config.rs:56-67)
    for ln in range(0, 5):
        out = oracle.hash_or_noop(x[:ln])
        assert list(out[:ln]) == list(x[:ln]) and all(int(v) == 0 for v in out[ln:])
    assert np.array_equal(oracle.hash_or_noop(x[:5]), oracle.hash_no_pad(x[:5]))
    # compress (hashing.rs:65-72)
    l, r = x[:4], x[4:8]
    st = np.zeros(12, dtype=np.uint64)
    st[:4], st[4:8] = l, r
    assert np.array_equal(oracle.two_to_one(l, r), oracle.poseidon(st)[:4])


@pytest.mark.parametrize("cap_height", [0, 1, 3, 8])
def test_merkle_all_proofs_verify(cap_height):
    # plonky2/src/hash/merkle_tree.rs:456-515 (verify_all_leaves; cap_height == log n edge)
    rng = np.random.default_rng(6)
    log_n = 8
    leaves = rand_field(rng, (1 << log_n, 7))
    digests, cap = oracle.merkle_tree(leaves, cap_height)
    assert digests.shape[0] == 2 * ((1 << log_n) - (1 << cap_height))
    for i in range(1 << log_n):
        sib = oracle.merkle_prove(digests, 1 << log_n, cap_height, i)
        assert oracle.merkle_verify(leaves[i], i, cap, sib)
    # tampering is detected
    bad = leaves[5].copy()
    bad[0] ^= np.uint64(1)
    assert not oracle.merkle_verify(bad, 5, cap, oracle.merkle_prove(digests, 1 << log_n, cap_height, 5))


def test_merkle_cap_height_too_big():
    # merkle_tree.rs:472-484 (should_panic)
    rng = np.random.default_rng(7)
    with pytest.raises(ValueError):
        oracle.merkle_tree(rand_field(rng, (256, 7)), 9)


def test_merkle_layout_closed_form():
    # node (layer l, position q) of a sub-tree lives at 2*(((q>>1) << (l+1)) + 2^l - 1) + (q&1)
    rng = np.random.default_rng(8)
    leaves = rand_field(rng, (64, 9))
    cap_height = 2
    digests, cap = oracle.merkle_tree(leaves, cap_height)
    sub_leaves = 64 >> cap_height
    sub_d = digests.shape[0] >> cap_height
    for t in range(1 << cap_height):
        layer = [oracle.hash_or_noop(leaves[t * sub_leaves + q]) for q in range(sub_leaves)]
        l = 0
        while len(layer) > 1:
            for q, h in enumerate(layer):
                idx = 2 * (((q >> 1) << (l + 1)) + (1 << l) - 1) + (q & 1)
                assert np.array_equal(digests[t * sub_d + idx], h)
            layer = [oracle.two_to_one(layer[2 * q], layer[2 * q + 1]) for q in range(len(layer) // 2)]
            l += 1
        assert np.array_equal(cap[t], layer[0])


@pytest.mark.parametrize("n_log,P_,rate_bits,cap_height,blinding", [(4, 3, 3, 4, False), (5, 9, 1, 0, True), (3, 5, 2, 5, False)])
def test_batch_matches_stepwise_definition(n_log, P_, rate_bits, cap_height, blinding):
    # PolynomialBatch::from_values == ifft -> lde/coset_fft -> transpose -> reverse_index_bits -> MerkleTree::new
    rng = np.random.default_rng(9)
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    values = rand_field(rng, (P_, n))
    salt = rand_field(rng, (4, N)) if blinding else None
    b = oracle.batch_from_values(values, rate_bits, cap_height, salt)
    cols = []
    for c in range(P_):
        co = oracle.ifft(values[c])
        assert np.array_equal(b.coeffs[c], co)
        assert np.array_equal(oracle.fft(co), values[c])
        cols.append(oracle.lde_coset_fft(co, rate_bits))
    if blinding:
        cols += [salt[i] for i in range(4)]
    lde = np.stack(cols)  # [P+salt][N]
    leaves = lde.T[[oracle.reverse_bits(L, n_log + rate_bits) for L in range(N)]]
    assert np.array_equal(b.leaves, leaves)
    digests, cap = oracle.merkle_tree(np.ascontiguousarray(leaves), cap_height)
    assert np.array_equal(b.digests, digests) and np.array_equal(b.cap, cap)
    # get_lde_values(i, step) strips the salt (oracle.rs:1007-1018)
    assert np.array_equal(b.get_lde_values(3, 1), lde[:P_, 3])
    # from_coeffs agrees
    b2 = oracle.batch_from_coeffs(b.coeffs, rate_bits, cap_height, salt)
    assert np.array_equal(b2.leaves, b.leaves) and np.array_equal(b2.cap, b.cap)
