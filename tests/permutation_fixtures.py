"""A witness with real copy constraints for the permutation-argument tests: routed wire cells are grouped into random
cycles, every cell of a cycle carries the same value and sigma maps each cell to the next one of its cycle
(sigma polynomial j at row i = k_{j'} * w^{i'}, plonk/permutation_argument.rs + circuit_builder.rs sigma_vecs)."""
import numpy as np

import oracle
from oracle.quotient import P


def make_permutation_instance(degree_bits, num_routed, num_wires, seed=0, honest=True):
    rng = np.random.default_rng(seed)
    n = 1 << degree_bits
    k_is = [pow(7, j, P) for j in range(num_routed)]
    w = oracle.primitive_root_of_unity(degree_bits)
    xs = [pow(w, i, P) for i in range(n)]
    cells = [(j, i) for j in range(num_routed) for i in range(n)]
    order = rng.permutation(len(cells))
    wires = rng.integers(0, P, size=(num_wires, n), dtype=np.uint64)
    sigma = np.zeros((num_routed, n), dtype=np.uint64)
    pos = 0
    while pos < len(order):
        ln = int(rng.integers(1, 6))
        cyc = [cells[t] for t in order[pos:pos + ln]]
        pos += ln
        v = rng.integers(0, P, dtype=np.uint64)
        for a, (j, i) in enumerate(cyc):
            wires[j, i] = v
            jn, inx = cyc[(a + 1) % len(cyc)]
            sigma[j, i] = k_is[jn] * xs[inx] % P
    if not honest:
        wires[0, 0] ^= np.uint64(1)
    return wires, sigma, k_is
