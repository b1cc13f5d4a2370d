"""CPU: pin the Python quotient oracle (oracle/quotient.py) with the reference's own property tests, restated:
  * each gate's constraints vanish on the witness its generator produces and not on a corrupted row
    (gates/gate_testing.rs::test_eval_fns spirit; the Poseidon gate row comes from the KAT-pinned permutation),
  * partial-products example (util/partial_products.rs:115-140),
  * an honest witness yields a quotient of degree < 7n (its top n coefficients vanish), a corrupted one does not:
    this exercises selector filters, the permutation argument, L_0, the alpha reduction and the Z_H division together.
"""
import numpy as np
import pytest

import oracle
from oracle import quotient as Q
from tests import quotient_fixtures as F

P = Q.P


def all_gates():
    (mix, _, _), _ = F.standard_gate_sets()
    return mix + [Q.PoseidonGate()]


@pytest.mark.parametrize("gate", all_gates(), ids=lambda g: type(g).__name__)
def test_gate_constraints_vanish_on_honest_rows(gate):
    rng = np.random.default_rng(5)
    for trial in range(4):
        consts = [F.rnd(rng) for _ in range(4)]
        pih = [F.rnd(rng) for _ in range(4)]
        w = F.honest_row(gate, rng, 135, consts, pih)
        cs = gate.eval_unfiltered(consts, w, pih)
        assert len(cs) == gate.num_constraints()
        assert all(c == 0 for c in cs), (type(gate).__name__, [i for i, c in enumerate(cs) if c])
    if gate.num_constraints():
        # flipping a wire the gate reads must break at least one constraint
        broke = 0
        for col in range(0, 30):
            w2 = list(w)
            w2[col] = (w2[col] + 1) % P
            broke += any(c != 0 for c in gate.eval_unfiltered(consts, w2, pih))
        assert broke > 0


def test_poseidon_gate_row_comes_from_the_pinned_permutation():
    g = Q.PoseidonGate()
    rng = np.random.default_rng(6)
    inp = [F.rnd(rng) for _ in range(12)]
    w = g.honest_row(inp, 0, 135)
    assert w[12:24] == [int(x) for x in oracle.poseidon(np.array(inp, dtype=np.uint64))]
    w = g.honest_row(inp, 1, 135)  # swap: first four inputs exchanged with the next four (Merkle sibling ordering)
    sw = inp[4:8] + inp[:4] + inp[8:]
    assert w[12:24] == [int(x) for x in oracle.poseidon(np.array(sw, dtype=np.uint64))]


def test_partial_products_reference_example():
    # util/partial_products.rs:115-140: v = 1..6, max_degree 2 -> partials [2, 24] with z_x = 1 and z_gx = 720
    nums, dens = [1, 2, 3, 4, 5, 6], [1] * 6
    assert Q.check_partial_products(nums, dens, [2, 24], 1, 720, 2) == [0, 0, 0]
    assert Q.check_partial_products(nums, dens, [2, 25], 1, 720, 2) != [0, 0, 0]
    assert -(-6 // 2) - 1 == 2  # num_partial_products(6, 2)


def test_filter_is_zero_exactly_for_the_other_gates_of_the_group():
    # gates/gate.rs:261-268 with selectors.rs:37-112
    group = (2, 5)
    for row in range(2, 5):
        for s in list(range(2, 5)) + [Q.UNUSED_SELECTOR]:
            f = Q.compute_filter(row, group, s, True)
            assert (f != 0) == (s == row)


def _quotient(inst):
    def commit(vals):
        return oracle.batch_from_values(vals, inst.circ.rate_bits, 0, want_digests=False).leaves
    return Q.compute_quotient_polys(inst.circ, commit(inst.wires), commit(inst.zs_pp), commit(inst.consts_sigmas), inst.pih,
                                    inst.betas, inst.gammas, inst.alphas)


@pytest.mark.parametrize("which", [0, 1])
def test_honest_witness_gives_low_degree_quotient(which):
    gates, groups, sel = F.standard_gate_sets()[which]
    degree_bits = 5
    n = 1 << degree_bits
    inst = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=11 + which)
    vals, coeffs = _quotient(inst)
    for c in coeffs:
        assert c.size == 8 * n
        assert not np.any(c[7 * n:]), "honest witness: quotient must have degree < 7n"
        assert np.any(c[:7 * n])
    bad = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=11 + which, corrupt=True)
    _, coeffs_bad = _quotient(bad)
    assert any(np.any(c[7 * n:]) for c in coeffs_bad), "a violated constraint must show up as a high-degree quotient"
