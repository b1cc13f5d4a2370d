"""Generates tests/golden/oracle_golden.json from the CPU oracle (run in the build container: `python tests/golden/make_golden.py`;
add `--headline` to also compute the cap of the 2^20 x 135 headline matrix, ~5 minutes on 8 cores).

The reference ships no stored vectors for LDE matrices, Merkle caps, quotient values or FRI proofs, so these
are SELF-GENERATED anchors: the oracle that produces them is pinned to the reference by its known-answer vectors and by the
reference's own CUDA kernels (tests/test_ref_cuda_crosscheck.py).  They freeze the outputs across rounds, give the GPU tests
fixed values to reproduce, and are what a machine with the Rust toolchain would compare against the real prover once.
Inputs are regenerated from the seeds below, so the file stays small: it stores caps and SHA-256 digests of the arrays.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import fri as FR  # noqa: E402
from oracle import quotient as Q  # noqa: E402

P = oracle.ORDER
MASK = (1 << 64) - 1


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK
    return x ^ (x >> 31)


def synthetic(count, seed, first=0):
    """p2b_fill_synthetic: splitmix64(seed ^ splitmix64(index)), re-mixed until < p.  Vectorised."""
    idx = np.arange(first, first + count, dtype=np.uint64)

    def mix(x):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    with np.errstate(over="ignore"):
        v = mix(np.uint64(seed) ^ mix(idx))
        bad = v >= np.uint64(P)
        while bad.any():
            v[bad] = mix(v[bad])
            bad = v >= np.uint64(P)
    return v


COMMIT_CASES = [  # (n_log, polys, rate_bits, cap_height, seed, blinding)
    (4, 3, 3, 2, 1, False), (6, 20, 3, 4, 2, False), (8, 135, 3, 4, 3, False), (5, 7, 1, 0, 4, True), (10, 234, 3, 4, 5, False)]


def commit_entry(n_log, polys, rate_bits, cap_height, seed, blinding):
    rng = np.random.default_rng(seed)
    values = rng.integers(0, P, size=(polys, 1 << n_log), dtype=np.uint64)
    salt = rng.integers(0, P, size=(4, 1 << (n_log + rate_bits)), dtype=np.uint64) if blinding else None
    b = oracle.batch_from_values(values, rate_bits, cap_height, salt=salt)
    return {"case": [n_log, polys, rate_bits, cap_height, seed, blinding], "cap": [[int(x) for x in h] for h in b.cap],
            "coeffs_sha256": sha(b.coeffs), "leaves_sha256": sha(b.leaves), "digests_sha256": sha(b.digests)}


def quotient_entry(which, degree_bits, seed):
    from tests import quotient_fixtures as F
    sets = F.standard_gate_sets() + (F.recursion_gate_set(),)
    gates, groups, sel = sets[which]
    inst = F.build_instance(gates, groups, sel, degree_bits, 135, 80, seed=seed)
    c = inst.circ
    ow = oracle.batch_from_values(inst.wires, c.rate_bits, 0, want_digests=False).leaves
    oz = oracle.batch_from_values(inst.zs_pp, c.rate_bits, 0, want_digests=False).leaves
    oc = oracle.batch_from_values(inst.consts_sigmas, c.rate_bits, 0, want_digests=False).leaves
    vals, coeffs = Q.compute_quotient_polys(c, ow, oz, oc, inst.pih, inst.betas, inst.gammas, inst.alphas)
    return {"case": [which, degree_bits, seed], "values_sha256": sha(np.array(vals)), "coeffs_sha256": sha(np.array(coeffs))}


def fri_entry(**kw):
    from tests.fri_fixtures import make_instance
    oracles, batches, params, ch = make_instance(**kw)
    pr = FR.prove_openings(batches, oracles, ch, params)
    return {"case": kw, "alpha": list(pr.alpha), "betas": [list(b) for b in pr.betas], "pow_witness": pr.pow_witness,
            "query_indices": pr.query_indices, "final_poly": [list(c) for c in pr.final_poly],
            "caps_sha256": [sha(c) for c in pr.commit_phase_merkle_caps], "challenger_state_after": ch.sponge_state}


def main():
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json")
    old = json.load(open(out_path)) if os.path.exists(out_path) else {}
    g = {"commits": [commit_entry(*c) for c in COMMIT_CASES],
         "quotients": [quotient_entry(0, 4, 21), quotient_entry(1, 4, 22), quotient_entry(2, 4, 81)],
         "fri": [fri_entry(), fri_entry(degree_bits=8, rate_bits=3, cap_height=4, arity_bits=[4], pow_bits=10, queries=6, seed=5)],
         "headline": old.get("headline")}
    if "--headline" in sys.argv:
        n_log, polys = 20, 135
        values = synthetic(polys << n_log, 0x504C4F4E4B5932).reshape(polys, 1 << n_log)
        b = oracle.batch_from_values(values, 3, 4, want_leaves=False, want_digests=False)
        g["headline"] = {"workload": "2^20 x 135, rate 3, cap 4, p2b_fill_synthetic seed 0x504C4F4E4B5932 (bench.py)",
                         "cap": [[int(x) for x in h] for h in b.cap], "cap_word0_hex": "%016x" % int(b.cap[0][0])}
    json.dump(g, open(out_path, "w"), indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
