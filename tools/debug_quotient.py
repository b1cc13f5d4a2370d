import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, plonky2_gpu_b200 as p2b
from oracle import quotient as Q
from tests import quotient_fixtures as F
from tests.test_gpu_quotient import run_both
ctx = p2b.Context()
def trial(name, gates, groups, sel, nw=135, nr=80, nch=2, randomize=False, seed=1):
    inst = F.build_instance(gates, groups, sel, 3, nw, nr, seed=seed, num_challenges=nch)
    if randomize:
        rng = np.random.default_rng(seed)
        for m in (inst.wires, inst.zs_pp, inst.consts_sigmas):
            m[:] = rng.integers(0, Q.P, size=m.shape, dtype=np.uint64)
    vals, coeffs, evals, ecoeffs = run_both(ctx, inst)
    ok = all(np.array_equal(vals[c], evals[c]) for c in range(nch))
    okc = all(np.array_equal(coeffs[c], ecoeffs[c]) for c in range(nch))
    print("%-28s values %s coeffs %s" % (name, ok, okc), flush=True)
    if not ok:
        print("   gpu", [hex(int(x)) for x in vals[0][:3]], "\n   ora", [hex(int(x)) for x in evals[0][:3]])
mix = F.standard_gate_sets()[0][0]
trial("noop honest nr=8 nch=1", [Q.NoopGate()], [(0, 1)], [0], nr=8, nch=1)
trial("noop random nr=8 nch=1", [Q.NoopGate()], [(0, 1)], [0], nr=8, nch=1, randomize=True)
trial("noop random nr=8 nch=2", [Q.NoopGate()], [(0, 1)], [0], nr=8, nch=2, randomize=True)
trial("noop random nr=80 nch=2", [Q.NoopGate()], [(0, 1)], [0], nr=80, nch=2, randomize=True)
for g in mix[1:] + [Q.PoseidonGate()]:
    trial("noop+" + type(g).__name__, [Q.NoopGate(), g], [(0, 2)], [0, 0], randomize=True)
trial("two groups", [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate()], [(0, 2), (2, 3)], [0, 0, 1], randomize=True)
