"""GPU: parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's KATs.
Bit-exact: everything is integer arithmetic mod p = 2^64 - 2^32 + 1."""
import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from tests.golden.reference_kats import POSEIDON_KATS

pytestmark = pytest.mark.gpu
P = oracle.ORDER


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


def rand_field(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64, endpoint=False).astype(np.uint64)


EDGE = np.array([0, 1, 2, P - 1, P - 2, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63,
                 0xFFFFFFFF00000000, 0xFFFFFFFE00000001, 0x00000000FFFFFFFF, 0xFFFFFFFFFFFFFFFE,
                 # products of these hit the rarely-taken repayment of reduce128 (lo + hi_lo*eps < hi_hi): 2^48 * 2^48 = 2^96
                 2**48, 3 * 2**48, 2**49 + 2**48, 0xFFFF * 2**48], dtype=np.uint64)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "mul_add", "add_canonical", "sub_canonical", "add_optimistic", "sub_optimistic"])
def test_field_ops_edge_and_random(ctx, op):
    rng = np.random.default_rng(11)
    a = np.concatenate([np.repeat(EDGE, EDGE.size), rng.integers(0, 2**64, size=4096, dtype=np.uint64)])
    b = np.concatenate([np.tile(EDGE, EDGE.size), rng.integers(0, 2**64, size=4096, dtype=np.uint64)])
    got = ctx.field_op(op, a, b)
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    if op in ("add", "add_canonical", "add_optimistic"):
        exp = [(x + y) % P for x, y in zip(ai, bi)]
    elif op in ("sub", "sub_canonical", "sub_optimistic"):
        exp = [(x - y) % P for x, y in zip(ai, bi)]
    elif op == "mul":
        exp = [(x * y) % P for x, y in zip(ai, bi)]
    else:
        exp = [(x * y + x) % P for x, y in zip(ai, bi)]
    assert [int(x) for x in got] == exp


def test_optimistic_add_sub_flag_exactly_the_second_wrap(ctx):
    """The NTT butterflies' add / sub repay the 64-bit wrap once and flag a second wrap (gl64.cuh): the flag must be set exactly
    when a + b - 2^64 >= p (add) or a - b + 2^64 < eps with a < b (sub), so an unflagged result is always a valid representative."""
    rng = np.random.default_rng(21)
    top = (2**64 - 1 - rng.integers(0, 2**33, size=4096, dtype=np.uint64)).astype(np.uint64)   # operands within 2^33 of 2^64
    low = rng.integers(0, 2**33, size=4096, dtype=np.uint64)
    a = np.concatenate([np.repeat(EDGE, EDGE.size), top, low, top])
    b = np.concatenate([np.tile(EDGE, EDGE.size), np.roll(top, 1), top, low])
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    eps = 2**32 - 1
    want_add = [1 if (x + y >= 2**64 and x + y - 2**64 + eps >= 2**64) else 0 for x, y in zip(ai, bi)]
    want_sub = [1 if (x < y and x - y + 2**64 < eps) else 0 for x, y in zip(ai, bi)]
    assert [int(x) for x in ctx.field_op("add_optimistic_flag", a, b)] == want_add
    assert [int(x) for x in ctx.field_op("sub_optimistic_flag", a, b)] == want_sub
    assert sum(want_add) > 100 and sum(want_sub) > 100      # the cases are actually exercised
    assert [int(x) for x in ctx.field_op("add_optimistic", a, b)] == [(x + y) % P for x, y in zip(ai, bi)]
    assert [int(x) for x in ctx.field_op("sub_optimistic", a, b)] == [(x - y) % P for x, y in zip(ai, bi)]


def test_poseidon_reference_known_answers(ctx):
    # plonky2/src/hash/poseidon_goldilocks.rs:277-318
    inp = np.array([k[0] for k in POSEIDON_KATS], dtype=np.uint64)
    out = ctx.poseidon(inp)
    for row, (_, exp) in zip(out, POSEIDON_KATS):
        assert [int(x) for x in row] == exp


def test_poseidon_random_and_noncanonical(ctx):
    rng = np.random.default_rng(12)
    s = rng.integers(0, 2**64, size=(2000, 12), dtype=np.uint64)  # includes non-canonical representatives
    s[:EDGE.size] = EDGE[:, None]
    got = ctx.poseidon(s)
    for i in range(0, 2000, 7):
        assert np.array_equal(got[i], oracle.poseidon(s[i])), i


@pytest.mark.parametrize("n_log", [0, 1, 2, 3, 5, 9, 10, 12, 13, 16])
def test_ifft_matches_oracle(ctx, n_log):
    rng = np.random.default_rng(13 + n_log)
    Pn = 5 if n_log < 14 else 3
    v = rand_field(rng, (Pn, 1 << n_log))
    got = ctx.ifft(v)
    for c in range(Pn):
        assert np.array_equal(got[c], oracle.ifft(v[c])), (n_log, c)


def test_ifft_large_two_strided_passes(ctx):
    rng = np.random.default_rng(14)
    n_log = 21  # 9 final + 12 -> two strided passes
    v = rand_field(rng, (2, 1 << n_log))
    got = ctx.ifft(v)
    assert np.array_equal(got[1], oracle.ifft(v[1]))
    assert np.array_equal(got[0], oracle.ifft(v[0]))


@pytest.mark.parametrize("n_log,rate_bits,Pn", [(0, 1, 3), (1, 3, 4), (3, 0, 9), (4, 3, 17), (9, 3, 9), (10, 2, 8), (12, 3, 19), (14, 1, 5)])
def test_lde_leaves_match_oracle(ctx, n_log, rate_bits, Pn):
    rng = np.random.default_rng(15)
    c = rand_field(rng, (Pn, 1 << n_log))
    got = ctx.lde_leaves(c, rate_bits)
    exp = oracle.batch_from_coeffs(c, rate_bits, 0, want_digests=False).leaves
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("leaf_len", [1, 2, 3, 4, 5, 7, 8, 9, 16, 20, 135, 234])
def test_merkle_tree_leaf_widths(ctx, leaf_len):
    # hash_or_noop for <= 4 elements, partial last sponge chunk, widths of the reference's presets
    rng = np.random.default_rng(16 + leaf_len)
    leaves = rng.integers(0, 2**64, size=(64, leaf_len), dtype=np.uint64)  # non-canonical inputs allowed
    for cap_height in (0, 2, 6):
        d, cap = ctx.merkle_tree(leaves, cap_height)
        ed, ecap = oracle.merkle_tree(leaves, cap_height)
        assert np.array_equal(cap, ecap), (leaf_len, cap_height)
        assert np.array_equal(d, ed), (leaf_len, cap_height)


def test_merkle_cap_height_too_big(ctx):
    # merkle_tree.rs:472-484 (should_panic) -> P2B_ERR_INVALID with the reference's message
    leaves = np.zeros((256, 7), dtype=np.uint64)
    with pytest.raises(p2b.P2BError, match="cap_height=9 should be at most"):
        ctx.merkle_tree(leaves, 9)


@pytest.mark.parametrize("n_log,Pn,rate_bits,cap_height,blinding", [
    (3, 3, 3, 4, False), (5, 9, 1, 0, True), (3, 5, 2, 5, False), (10, 20, 3, 4, False),
    (12, 135, 3, 4, False), (13, 16, 3, 4, True), (9, 2, 3, 12, False), (2, 1, 0, 0, False)])
def test_commit_from_values_matches_oracle(ctx, n_log, Pn, rate_bits, cap_height, blinding):
    rng = np.random.default_rng(17)
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    values = rand_field(rng, (Pn, n))
    salt = rand_field(rng, (4, N)) if blinding else None
    b = p2b.PolynomialBatch.from_values(ctx, values, rate_bits, cap_height, blinding=blinding, salt=salt)
    e = oracle.batch_from_values(values, rate_bits, cap_height, salt)
    assert np.array_equal(b.polynomials(), e.coeffs)
    assert np.array_equal(b.cap(), e.cap)
    assert np.array_equal(b.digests(), e.digests)
    assert np.array_equal(b.leaves(), e.leaves)
    for idx in (0, 1, N // 2, N - 1):
        assert np.array_equal(b.get_lde_values(idx, 1), e.get_lde_values(idx, 1))
    # from_coeffs on the same coefficients gives the same commitment
    b2 = p2b.PolynomialBatch.from_coeffs(ctx, e.coeffs, rate_bits, cap_height, blinding=blinding, salt=salt)
    assert np.array_equal(b2.cap(), e.cap) and np.array_equal(b2.leaves(), e.leaves)
    # every opened row verifies against the cap with the CPU verifier (merkle_proofs.rs:53-81)
    idxs = sorted(set([0, 1, N - 1] + [int(x) for x in rng.integers(0, N, size=8)]))
    rows, sibs = b.open_rows(idxs)
    for r, s, i in zip(rows, sibs, idxs):
        assert np.array_equal(r, e.leaves[i])
        assert np.array_equal(s, oracle.merkle_prove(e.digests, N, cap_height, i))
        assert oracle.merkle_verify(r, i, b.cap(), s)
    b.close()
    b2.close()


def test_commit_device_resident_input(ctx):
    rng = np.random.default_rng(18)
    values = rand_field(rng, (7, 1 << 11))
    d = p2b.DeviceBuffer.from_host(ctx, values)
    b = p2b.PolynomialBatch.from_values(ctx, (d, 7, 1 << 11), 3, 4)
    e = oracle.batch_from_values(values, 3, 4)
    assert np.array_equal(b.cap(), e.cap)
    assert np.array_equal(d.to_host().reshape(7, -1), values)  # caller's buffer is left untouched
    b.close()


def test_commit_error_paths(ctx):
    with pytest.raises(p2b.P2BError):
        p2b.PolynomialBatch.from_values(ctx, np.zeros((0, 8), dtype=np.uint64), 3, 4)
    with pytest.raises(p2b.P2BError, match="cap_height"):
        p2b.PolynomialBatch.from_values(ctx, np.zeros((2, 8), dtype=np.uint64), 1, 5)
    with pytest.raises(p2b.P2BError):
        p2b.PolynomialBatch.from_values(ctx, np.zeros((2, 12), dtype=np.uint64), 1, 0)


def test_commit_linearity_property(ctx):
    # size-independent property: LDE rows are linear in the input values
    rng = np.random.default_rng(19)
    n_log, Pn = 14, 6
    a, b = rand_field(rng, (Pn, 1 << n_log)), rand_field(rng, (Pn, 1 << n_log))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ba = p2b.PolynomialBatch.from_values(ctx, a, 3, 4)
    bb = p2b.PolynomialBatch.from_values(ctx, b, 3, 4)
    bs = p2b.PolynomialBatch.from_values(ctx, s, 3, 4)
    rows = [0, 5, 1 << 15, (1 << 17) - 1]
    ra, rb, rs = ba.open_rows(rows, False)[0], bb.open_rows(rows, False)[0], bs.open_rows(rows, False)[0]
    assert np.array_equal(((ra.astype(object) + rb.astype(object)) % P).astype(np.uint64), rs)
    for x in (ba, bb, bs):
        x.close()


def test_commit_host_pipeline_with_coefficient_copy_back(ctx):
    """p2b_commit_from_values_ex: grouped H2D overlapped with the inverse NTT, coefficients copied back to pinned host
    memory while the tree is built (the data movement of fri/oracle.rs:352-362, 403-407)."""
    rng = np.random.default_rng(23)
    for n_log, Pn in [(11, 37), (8, 16), (13, 135)]:
        n = 1 << n_log
        values = rand_field(rng, (Pn, n))
        pinned_in, pinned_out = p2b.PinnedBuffer(Pn * n), p2b.PinnedBuffer(Pn * n)
        pinned_in.array[:] = values.reshape(-1)
        hv, hc = pinned_in.array.reshape(Pn, n), pinned_out.array.reshape(Pn, n)
        b = p2b.PolynomialBatch.from_values(ctx, hv, 3, 4, coeffs_out=hc)
        e = oracle.batch_from_values(values, 3, 4)
        assert np.array_equal(b.cap(), e.cap)           # synchronises the stream, which also covers the copy
        assert np.array_equal(hc, e.coeffs)
        assert np.array_equal(b.polynomials(), e.coeffs)
        assert np.array_equal(b.digests(), e.digests)
        b.close()
        pinned_in.free()
        pinned_out.free()


def test_poseidon_optimistic_reduction_falls_back_exactly(ctx):
    """The hash kernels run an optimistic reduce128 (gl64.cuh) and redo a leaf / node / state exactly when the rare
    borrow case fires.  Force it: a state lane equal to 2^48 - RC[0][lane] makes the first S-box square 2^48 (x*x = 2^96:
    lo = 0, hi = 2^32), for the permutation entry, a leaf hash and a 2-to-1 compression."""
    from oracle import poseidon_params as PP
    rng = np.random.default_rng(29)
    states = rng.integers(0, P, size=(64, 12), dtype=np.uint64)
    for k in range(64):
        lane = k % 12
        states[k, lane] = (2**48 * (1 + k // 12) - PP.ROUND_CONSTANTS[lane]) % P
    got = ctx.poseidon(states)
    for k in range(64):
        assert np.array_equal(got[k], oracle.poseidon(states[k])), k
    # leaves whose first absorbed chunk (lanes 0..7) triggers the case, in every 3rd leaf; capacity lanes start at 0,
    # so lanes 8..11 would need RC = 2^48 and are left alone
    leaves = rng.integers(0, P, size=(128, 20), dtype=np.uint64)
    for r in range(0, 128, 3):
        lane = r % 8
        leaves[r, lane] = (2**48 - PP.ROUND_CONSTANTS[lane]) % P
    d, cap = ctx.merkle_tree(leaves, 2)
    ed, ecap = oracle.merkle_tree(leaves, 2)
    assert np.array_equal(d, ed) and np.array_equal(cap, ecap)
    # 4-element leaves are copied verbatim (hash_or_noop), so layer 1 compresses exactly these crafted values
    small = rng.integers(0, P, size=(64, 4), dtype=np.uint64)
    for r in range(0, 64, 2):
        small[r, r % 4] = (2**48 - PP.ROUND_CONSTANTS[r % 4]) % P
    d, cap = ctx.merkle_tree(small, 0)
    ed, ecap = oracle.merkle_tree(small, 0)
    assert np.array_equal(d, ed) and np.array_equal(cap, ecap)


def test_ntt_exact_redo_path(ctx):
    """NTT tiles run optimistic butterflies and are redone exactly when a reduction reports its rare case; the test hook
    forces that path for every tile (strided, final-rows) and the results must not change."""
    rng = np.random.default_rng(31)
    values = rand_field(rng, (9, 1 << 12))
    ctx.debug_force_exact_redo(True)
    try:
        b = p2b.PolynomialBatch.from_values(ctx, values, 3, 4)
        e = oracle.batch_from_values(values, 3, 4)
        assert np.array_equal(b.polynomials(), e.coeffs)
        assert np.array_equal(b.leaves(), e.leaves)
        assert np.array_equal(b.cap(), e.cap)
        b.close()
    finally:
        ctx.debug_force_exact_redo(False)
