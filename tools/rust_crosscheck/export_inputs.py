"""Writes the inputs of every cross-check case as flat little-endian u64 files (+ a .json with shapes) for
tools/rust_crosscheck/crosscheck.rs.  Same seeds / fixtures as tests/golden/make_golden.py and tests/*_fixtures.py.
    python tools/rust_crosscheck/export_inputs.py <out_dir>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from oracle import recursion_gates as RG  # noqa: E402
from tests import quotient_fixtures as QF  # noqa: E402
from tests.fri_fixtures import make_instance  # noqa: E402
from tests.golden.make_golden import COMMIT_CASES  # noqa: E402
from tests.permutation_fixtures import make_permutation_instance  # noqa: E402

P = oracle.ORDER
FRI_CASES = [dict(), dict(degree_bits=8, rate_bits=3, cap_height=4, arity_bits=[4], pow_bits=10, queries=6, seed=5)]
PERM_CASES = [(4, 11, 14, 3, [5, 6], [7, 8], 4), (6, 80, 135, 9, [11, 12], [13, 14], 8)]  # degree_bits, routed, wires, seed, betas, gammas, max_degree
GATES = [("arithmetic_extension", RG.ArithmeticExtensionGate(10)), ("mul_extension", RG.MulExtensionGate(13)),
         ("reducing", RG.ReducingGate(43)), ("reducing_extension", RG.ReducingExtensionGate(32)),
         ("exponentiation", RG.ExponentiationGate(66)), ("poseidon_mds", RG.PoseidonMdsGate()),
         ("high_degree_interpolation", RG.HighDegreeInterpolationGate(2)), ("low_degree_interpolation", RG.LowDegreeInterpolationGate(4))]


def w(out, name, arr):
    np.ascontiguousarray(np.asarray(arr, dtype=np.uint64)).tofile(os.path.join(out, name + ".bin"))


def commit_inputs(case):
    n_log, polys, rate_bits, cap_height, seed, blinding = case
    rng = np.random.default_rng(seed)
    return rng.integers(0, P, size=(polys, 1 << n_log), dtype=np.uint64)


def gate_rows(gate, nw=135, ncst=4, rows=16, seed=5):
    rng = np.random.default_rng(seed)
    pih = [int(x) for x in rng.integers(0, P, size=4, dtype=np.uint64)]
    wires = np.zeros((rows, nw), dtype=np.uint64)
    consts = np.zeros((rows, ncst), dtype=np.uint64)
    for r in range(rows):
        gc = [int(x) for x in rng.integers(0, P, size=ncst, dtype=np.uint64)]
        wires[r] = np.array(QF.honest_row(gate, rng, nw, gc, pih), dtype=np.uint64) if r % 2 == 0 else rng.integers(0, P, size=nw, dtype=np.uint64)
        consts[r] = np.array(gc, dtype=np.uint64)
    return wires, consts, pih


def main(out):
    os.makedirs(out, exist_ok=True)
    meta = {}
    k = 0
    for case in COMMIT_CASES:
        if case[5]:
            continue  # blinding draws its salt from the reference's own RNG: not reproducible across implementations
        w(out, "commit_%d.meta" % k, case[:4])
        w(out, "commit_%d.values" % k, commit_inputs(case))
        meta["commit_%d" % k] = list(case)
        k += 1
    for k, kw in enumerate(FRI_CASES):
        oracles, batches, params, ch = make_instance(**kw)
        polys = [o.values.shape[0] for o in oracles]
        w(out, "fri_%d.meta" % k, [params.degree_bits, params.rate_bits, params.cap_height, params.proof_of_work_bits, params.num_query_rounds,
                                   len(params.reduction_arity_bits)] + list(params.reduction_arity_bits) + [len(oracles)] + polys)
        for j, o in enumerate(oracles):
            w(out, "fri_%d.oracle%d.values" % (k, j), o.values)
        w(out, "fri_%d.zeta" % k, list(batches[0].point))
        rng = np.random.default_rng(kw.get("seed", 1))     # replay make_instance's draws up to the transcript prefix
        n = 1 << params.degree_bits
        for kk in polys:
            rng.integers(0, P, size=(kk, n), dtype=np.uint64)
        rng.integers(0, P, dtype=np.uint64), rng.integers(0, P, dtype=np.uint64)
        w(out, "fri_%d.observed" % k, rng.integers(0, P, size=11, dtype=np.uint64))
        meta["fri_%d" % k] = kw
    for k, (db, routed, nw, seed, betas, gammas, md) in enumerate(PERM_CASES):
        wires, sigma, k_is = make_permutation_instance(db, routed, nw, seed=seed)
        w(out, "perm_%d.meta" % k, [db, routed, len(betas), md])
        w(out, "perm_%d.wires" % k, wires[:routed])
        w(out, "perm_%d.sigmas" % k, sigma)
        w(out, "perm_%d.k_is" % k, k_is)
        w(out, "perm_%d.betas" % k, betas)
        w(out, "perm_%d.gammas" % k, gammas)
        meta["perm_%d" % k] = [db, routed, nw, seed, betas, gammas, md]
    for name, gate in GATES:
        wires, consts, pih = gate_rows(gate)
        w(out, "gate_%s.meta" % name, [wires.shape[0], wires.shape[1], consts.shape[1]])
        w(out, "gate_%s.wires" % name, wires)
        w(out, "gate_%s.consts" % name, consts)
        w(out, "gate_%s.pih" % name, pih)
        meta["gate_" + name] = list(gate.params) if hasattr(gate, "params") else []
    json.dump(meta, open(os.path.join(out, "cases.json"), "w"), indent=1)
    print("wrote", len(os.listdir(out)), "files to", out)


if __name__ == "__main__":
    main(sys.argv[1])
