"""CPU: the recursion gate set's constraint restatement (oracle/recursion_gates.py), checked the way the reference's
gate_testing.rs + per-gate tests do: constraints vanish on the generator's row, a corrupted wire breaks them, and the
constraint counts match each gate's num_constraints()."""
import numpy as np
import pytest

from oracle import recursion_gates as RG
from oracle.quotient import P


def gates():
    return [RG.ArithmeticExtensionGate(10), RG.MulExtensionGate(13), RG.ReducingGate(43), RG.ReducingExtensionGate(32),
            RG.ExponentiationGate(66), RG.PoseidonMdsGate(), RG.HighDegreeInterpolationGate(2), RG.HighDegreeInterpolationGate(4),
            RG.LowDegreeInterpolationGate(2), RG.LowDegreeInterpolationGate(4), RG.LowDegreeInterpolationGate(1)]


@pytest.mark.parametrize("gate", gates(), ids=lambda g: type(g).__name__ + str(g.params))
def test_constraints_vanish_on_honest_rows_only(gate):
    rng = np.random.default_rng(7)
    rnd = lambda: int(rng.integers(0, P, dtype=np.uint64))
    consts = [rnd(), rnd()]
    w = gate.honest_row(rnd, 135, consts)
    cs = gate.eval_unfiltered(consts, w, [0] * 4)
    assert len(cs) == gate.num_constraints()
    assert all(c == 0 for c in cs)
    w2 = list(w)
    w2[0] = (w2[0] + 1) % P          # wire 0 is read by every one of these gates
    assert any(c != 0 for c in gate.eval_unfiltered(consts, w2, [0] * 4))


def test_standard_recursion_config_shapes():
    # constraint counts of the gates as standard_recursion_config instantiates them (135 wires, 80 routed)
    assert RG.ArithmeticExtensionGate(RG.ArithmeticExtensionGate.num_ops_for(80)).num_constraints() == 20
    assert RG.ReducingGate(43).min_wires() <= 135 and RG.ReducingGate(43).num_constraints() == 86
    assert RG.ReducingExtensionGate(32).min_wires() <= 135
    assert RG.ExponentiationGate(66).num_constraints() == 67          # max_power_bits(135, 80) = min(78, 66)
    assert RG.LowDegreeInterpolationGate(4).min_wires() <= 135 and RG.LowDegreeInterpolationGate(4).num_constraints() == 14 + 32 + 28 + 2
    assert RG.HighDegreeInterpolationGate(4).num_constraints() == 34
    assert RG.PoseidonMdsGate().num_constraints() == 24


def test_exponentiation_gate_computes_powers():
    g = RG.ExponentiationGate(5)
    rng = np.random.default_rng(1)
    w = g.honest_row(lambda: int(rng.integers(0, P, dtype=np.uint64)), 20, [])
    e = sum(b << i for i, b in enumerate(w[1:6]))
    assert w[6] == pow(w[0], e, P)
