"""CPU: restatement of wires_permutation_partial_products_and_zs (plonk/prover.rs:729-786) -- the properties the
protocol relies on: Z(1) = 1, the partial-product recurrence checked by check_partial_products
(util/partial_products.rs:52-78) vanishes on H, and Z wraps around to 1 exactly when the copy constraints hold."""
import numpy as np

import oracle
from oracle import quotient as Q
from oracle.quotient import P
from tests.permutation_fixtures import make_permutation_instance


def recurrence_residuals(wires, sigma, k_is, beta, gamma, deg, degree_bits, cols):
    n = 1 << degree_bits
    nr = len(k_is)
    w = oracle.primitive_root_of_unity(degree_bits)
    pps, z = cols[:-1], cols[-1]
    out = []
    for i in range(n):
        x = pow(w, i, P)
        nums = [(int(wires[j][i]) + beta * (k_is[j] * x % P) + gamma) % P for j in range(nr)]
        dens = [(int(wires[j][i]) + beta * int(sigma[j][i]) + gamma) % P for j in range(nr)]
        out += Q.check_partial_products(nums, dens, [pp[i] for pp in pps], z[i], z[(i + 1) % n], deg)
    return out


def test_z_and_partial_products_satisfy_the_recurrence():
    db, nr, deg = 4, 11, 4   # 3 chunks, the last one short
    wires, sigma, k_is = make_permutation_instance(db, nr, 14, seed=3)
    beta, gamma = 0x1234567890abcdef % P, 0xfedcba0987654321 % P
    cols = Q.wires_permutation_partial_products_and_zs(wires, sigma, k_is, beta, gamma, deg, db)
    assert len(cols) == -(-nr // deg) and cols[-1][0] == 1
    assert all(r == 0 for r in recurrence_residuals(wires, sigma, k_is, beta, gamma, deg, db, cols))


def test_z_wraps_to_one_iff_copy_constraints_hold():
    db, nr, deg = 3, 8, 3
    beta, gamma = 77777777777, 99999999999
    for honest in (True, False):
        wires, sigma, k_is = make_permutation_instance(db, nr, 9, seed=4, honest=honest)
        cols = Q.wires_permutation_partial_products_and_zs(wires, sigma, k_is, beta, gamma, deg, db)
        res = recurrence_residuals(wires, sigma, k_is, beta, gamma, deg, db, cols)
        # every residual vanishes except (when cheating) the wrap-around Z(g^n) = Z(1) of the last row
        assert all(r == 0 for r in res[:-1])
        assert (res[-1] == 0) == honest


def test_matrix_order_is_zs_first():
    db, nr, deg = 2, 6, 2
    wires, sigma, k_is = make_permutation_instance(db, nr, 6, seed=5)
    betas, gammas = [3, 5], [7, 11]
    m = Q.all_zs_partial_products(wires, sigma, k_is, betas, gammas, deg, db)
    per0 = Q.wires_permutation_partial_products_and_zs(wires, sigma, k_is, 3, 7, deg, db)
    per1 = Q.wires_permutation_partial_products_and_zs(wires, sigma, k_is, 5, 11, deg, db)
    assert m[0] == per0[-1] and m[1] == per1[-1]
    assert m[2:4] == per0[:-1] and m[4:6] == per1[:-1]
