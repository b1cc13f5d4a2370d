"""Compares what the reference wrote (crosscheck.rs) with the CPU oracle, bit for bit.
    python tools/rust_crosscheck/compare.py <in_dir> <out_dir>
    python tools/rust_crosscheck/compare.py <in_dir> --self-test     (writes the ORACLE's outputs in the reference's format
                                                                       into <in_dir>/_self and compares: checks this script)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from oracle import fri as FR  # noqa: E402
from oracle import quotient as Q  # noqa: E402
import export_inputs as X  # noqa: E402
from tests.fri_fixtures import make_instance  # noqa: E402


def expected():
    """{file stem: uint64 array} in the layout crosscheck.rs writes."""
    out = {}
    k = 0
    for case in X.COMMIT_CASES:
        if case[5]:
            continue
        b = oracle.batch_from_values(X.commit_inputs(case), case[2], case[3])
        out["commit_%d.coeffs" % k], out["commit_%d.leaves" % k] = b.coeffs, b.leaves
        out["commit_%d.digests" % k], out["commit_%d.cap" % k] = b.digests, b.cap
        k += 1
    for k, kw in enumerate(X.FRI_CASES):
        oracles, batches, params, ch = make_instance(**kw)
        pr = FR.prove_openings(batches, oracles, ch, params)
        for i, cap in enumerate(pr.commit_phase_merkle_caps):
            out["fri_%d.cap%d" % (k, i)] = cap
        out["fri_%d.final_poly" % k] = np.array(pr.final_poly, dtype=np.uint64)
        out["fri_%d.pow_witness" % k] = np.array([pr.pow_witness], dtype=np.uint64)
        rows = []
        for initial, steps in pr.query_round_proofs:       # (initial [(row, siblings)], steps [(evals flattened, siblings)])
            for leaf, path in initial:
                rows += [int(x) for x in leaf] + [int(x) for h in path for x in h]
            for evals, path in steps:
                rows += [int(x) for x in np.asarray(evals).reshape(-1)] + [int(x) for h in path for x in h]
        out["fri_%d.queries" % k] = np.array(rows, dtype=np.uint64)
        c2 = ch.clone()                                   # Challenger::compact (iop/challenger.rs): absorb what is pending, return the state
        if len(c2.input_buffer):
            c2.duplexing()
        out["fri_%d.challenger_after" % k] = np.array(c2.sponge_state, dtype=np.uint64)
    for k, (db, routed, nw, seed, betas, gammas, md) in enumerate(X.PERM_CASES):
        wires, sigma, k_is = X.make_permutation_instance(db, routed, nw, seed=seed)
        cols = []
        for b_, g_ in zip(betas, gammas):
            pp = Q.wires_permutation_partial_products_and_zs(wires, sigma, k_is, b_, g_, md, db)   # partial products.., Z last
            cols += [pp[-1]] + pp[:-1]                                                          # the dumper writes Z first
        out["perm_%d.zs_pp" % k] = np.array(cols, dtype=np.uint64)
    for name, gate in X.GATES:
        wires, consts, pih = X.gate_rows(gate)
        per_row = [gate.eval_unfiltered([int(x) for x in consts[r]], [int(x) for x in wires[r]], pih) for r in range(wires.shape[0])]
        out["gate_%s.constraints" % name] = np.array(per_row, dtype=np.uint64).T.copy()         # constraint-major, as eval_unfiltered_base_batch
    return out


def main():
    in_dir = sys.argv[1]
    exp = expected()
    if sys.argv[2] == "--self-test":
        out_dir = os.path.join(in_dir, "_self")
        os.makedirs(out_dir, exist_ok=True)
        for stem, arr in exp.items():
            np.ascontiguousarray(np.asarray(arr, dtype=np.uint64)).tofile(os.path.join(out_dir, stem + ".bin"))
    else:
        out_dir = sys.argv[2]
    bad = 0
    for stem, arr in sorted(exp.items()):
        path = os.path.join(out_dir, stem + ".bin")
        if not os.path.exists(path):
            print("MISSING  ", stem)
            bad += 1
            continue
        got = np.fromfile(path, dtype=np.uint64)
        want = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64)).reshape(-1)
        ok = got.shape == want.shape and np.array_equal(got, want)
        print("ok       " if ok else "MISMATCH ", stem, got.shape, want.shape)
        bad += 0 if ok else 1
    print("%d of %d outputs differ" % (bad, len(exp)) if bad else "all %d outputs identical to the oracle" % len(exp))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
