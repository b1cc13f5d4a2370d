"""Synthetic FRI opening instances shared by the CPU oracle tests and the GPU parity tests: four random committed
batches standing in for constants_sigmas / wires / zs_partial_products / quotient (plonk/plonk_common.rs FRI_ORACLES),
all opened at zeta and the third also at g*zeta (plonk/circuit_data.rs:351-371)."""
import numpy as np

import oracle
from oracle import fri as FR
from oracle.quotient import P


def make_instance(degree_bits=5, rate_bits=2, cap_height=1, polys=(3, 5, 2, 2), seed=1, salted=(False, False, False, False),
                  arity_bits=(2, 1), pow_bits=3, queries=4):
    rng = np.random.default_rng(seed)
    n = 1 << degree_bits
    oracles = []
    for k, s in zip(polys, salted):
        vals = rng.integers(0, P, size=(k, n), dtype=np.uint64)
        salt = rng.integers(0, P, size=(4, n << rate_bits), dtype=np.uint64) if s else None
        oracles.append(oracle.batch_from_values(vals, rate_bits, cap_height, salt=salt))
        oracles[-1].values, oracles[-1].salt = vals, salt
    zeta = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))
    g = oracle.primitive_root_of_unity(degree_bits)
    all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
    zs = [(2, p) for p in range(polys[2])]
    batches = [FR.FriBatchInfo(zeta, all_polys), FR.FriBatchInfo(FR.escale(zeta, g), zs)]
    params = FR.FriParams(degree_bits, rate_bits, cap_height, pow_bits, queries, arity_bits, hiding=any(salted))
    ch = FR.Challenger()
    ch.observe_elements([int(x) for x in rng.integers(0, P, size=11, dtype=np.uint64)])  # some earlier transcript
    return oracles, batches, params, ch
