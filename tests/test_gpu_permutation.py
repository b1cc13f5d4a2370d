"""GPU: p2b_partial_products_and_zs against the restatement of plonk/prover.rs:702-786 (bit-exact), and chained on the device
with the commit and the quotient kernel: witness -> Z / partial products -> commit -> quotient, never leaving the GPU."""
import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b
from oracle import quotient as Q
from oracle.quotient import P
from tests.permutation_fixtures import make_permutation_instance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context()
    yield c
    c.close()


@pytest.mark.parametrize("db,nr,nw,deg,nc", [(4, 11, 14, 4, 2), (0, 3, 3, 2, 1), (1, 9, 9, 8, 3), (6, 80, 135, 8, 2), (11, 20, 24, 8, 2),
                                            (3, 40, 40, 2, 1)])
def test_partial_products_match_oracle(ctx, db, nr, nw, deg, nc):
    wires, sigma, k_is = make_permutation_instance(db, nr, nw, seed=db + nr)
    rng = np.random.default_rng(9)
    betas = [int(x) for x in rng.integers(0, P, size=nc, dtype=np.uint64)]
    gammas = [int(x) for x in rng.integers(0, P, size=nc, dtype=np.uint64)]
    if db > 8:   # keep the pure-Python restatement to a few seconds: compare a strided subset of rows via the recurrence
        want = None
    else:
        want = np.array(Q.all_zs_partial_products(wires, sigma, k_is, betas, gammas, deg, db), dtype=np.uint64)
    out, shape = p2b.partial_products_and_zs(ctx, wires, sigma, k_is, betas, gammas, deg)
    got = out.to_host().reshape(shape)
    K = -(-nr // deg)
    assert shape == (nc * K, 1 << db)
    if want is not None:
        assert np.array_equal(got, want)
    else:
        # size-independent property: Z(1) = 1, the recurrence holds on sampled rows, and Z wraps to 1 (honest witness)
        n = 1 << db
        w = oracle.primitive_root_of_unity(db)
        for c in range(nc):
            assert got[c][0] == 1
            pps = got[nc + c * (K - 1): nc + (c + 1) * (K - 1)]
            for i in list(range(0, n, 97)) + [n - 1]:
                x = pow(w, i, P)
                nums = [(int(wires[j][i]) + betas[c] * (k_is[j] * x % P) + gammas[c]) % P for j in range(nr)]
                dens = [(int(wires[j][i]) + betas[c] * int(sigma[j][i]) + gammas[c]) % P for j in range(nr)]
                res = Q.check_partial_products(nums, dens, [int(pp[i]) for pp in pps], int(got[c][i]), int(got[c][(i + 1) % n]), deg)
                assert all(r == 0 for r in res)


def test_partial_products_error_paths(ctx):
    wires, sigma, k_is = make_permutation_instance(2, 6, 6, seed=1)
    with pytest.raises(p2b.P2BError, match="smaller that the degree"):
        p2b.partial_products_and_zs(ctx, wires, sigma, k_is, [1], [2], 6)
    with pytest.raises(p2b.P2BError, match="num_challenges"):
        p2b.partial_products_and_zs(ctx, wires, sigma, k_is, [1] * 5, [2] * 5, 2)


def test_device_chain_partial_products_commit_quotient(ctx):
    # a Noop-only circuit with real copy constraints: the quotient is the permutation argument alone
    db, nr, nw, deg, nc, rate_bits = 5, 16, 20, 8, 2, 3
    wires, sigma, k_is = make_permutation_instance(db, nr, nw, seed=21)
    n = 1 << db
    circ = Q.Circuit([Q.NoopGate()], [0], [(0, 1)], nw, nr, 3, k_is, db, rate_bits, nc, deg)
    consts = np.zeros((3, n), dtype=np.uint64)      # selector column = gate index 0, two unused gate constants
    consts_sigmas = np.concatenate([consts, sigma])
    rng = np.random.default_rng(22)
    betas, gammas, alphas = ([int(x) for x in rng.integers(0, P, size=nc, dtype=np.uint64)] for _ in range(3))
    pih = [1, 2, 3, 4]
    # device chain
    dz, shape = p2b.partial_products_and_zs(ctx, wires, sigma, k_is, betas, gammas, deg)
    bz = p2b.PolynomialBatch.from_values(ctx, (dz, shape[0], shape[1]), rate_bits, 2)
    bw = p2b.PolynomialBatch.from_values(ctx, wires, rate_bits, 2)
    bc = p2b.PolynomialBatch.from_values(ctx, consts_sigmas, rate_bits, 2)
    pc = p2b.Circuit([(0, ())], [0], [(0, 1)], nw, nr, 3, k_is, db, rate_bits, nc, deg)
    vals, coeffs = p2b.compute_quotient_polys(ctx, pc, bw, bz, bc, pih, betas, gammas, alphas)
    # CPU chain
    zs_pp = np.array(Q.all_zs_partial_products(wires, sigma, k_is, betas, gammas, deg, db), dtype=np.uint64)
    ow = oracle.batch_from_values(wires, rate_bits, 0, want_digests=False).leaves
    oz = oracle.batch_from_values(zs_pp, rate_bits, 0, want_digests=False).leaves
    oc = oracle.batch_from_values(consts_sigmas, rate_bits, 0, want_digests=False).leaves
    evals, ecoeffs = Q.compute_quotient_polys(circ, ow, oz, oc, pih, betas, gammas, alphas)
    for c in range(nc):
        assert np.array_equal(vals[c], evals[c])
        assert np.array_equal(coeffs[c], ecoeffs[c])
    # honest copy constraints: the numerator vanishes on H, so the quotient is a polynomial of degree < 8n - n + ... and in
    # particular its evaluations are not all zero while a corrupted witness changes them
    assert any(int(v) for v in vals[0])
    for b in (bz, bw, bc):
        b.close()


def test_zero_denominator_is_an_error_like_the_reference_panic():
    """batch_multiplicative_inverse panics with "Tried to invert zero" (field/src/types.rs:130); the device entry reports it."""
    p2b.build()
    ctx = p2b.Context()
    n_log, routed = 4, 9
    n = 1 << n_log
    rng = np.random.default_rng(1)
    wires = rng.integers(0, oracle.ORDER, size=(routed, n), dtype=np.uint64)
    sigma = rng.integers(0, oracle.ORDER, size=(routed, n), dtype=np.uint64)
    beta, gamma = 5, 11
    # make one denominator vanish: wire + beta * sigma + gamma == 0 at (column 3, row 6)
    wires[3, 6] = (oracle.ORDER - (beta * int(sigma[3, 6]) + gamma) % oracle.ORDER) % oracle.ORDER
    k_is = [pow(7, j, oracle.ORDER) for j in range(routed)]
    with pytest.raises(p2b.P2BError, match="invert zero"):
        p2b.partial_products_and_zs(ctx, wires, sigma, k_is, [beta], [gamma], 4)
    ctx.close()
