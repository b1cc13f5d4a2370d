"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's FRI opening proof (prover AND verifier).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product (plonky2-gpu_b200) never does.

Follows (reference file:line):
  * quadratic extension F[X]/(X^2 - 7):      field/src/extension/quadratic.rs:86-99,172-185, goldilocks_extensions.rs:14-28
  * Challenger (duplex sponge, overwrite):    plonky2/src/iop/challenger.rs:15-150
  * OpeningSet evaluation (Horner):           plonky2/src/plonk/proof.rs:305-334, field/src/polynomial/mod.rs:161-166
  * ReducingFactor / divide_by_linear:        plonky2/src/util/reducing.rs:25-110, field/src/polynomial/division.rs:73-88
  * prove_openings:                           plonky2/src/fri/oracle.rs:1046-1110
  * fri_proof / commit phase / PoW / queries: plonky2/src/fri/prover.rs:23-260
  * fri_challenges:                           plonky2/src/fri/challenges.rs:24-66
  * verifier:                                 plonky2/src/fri/verifier.rs:18-260, hash/merkle_proofs.rs:53-79

Parity pinning: the reference holds no golden vector for a FRI proof.  This restatement is pinned (a) transitively -- the
Poseidon permutation, NTT and Merkle tree it calls are the KAT-checked C oracle -- and (b) by the verifier restated below
accepting the prover's output and rejecting corrupted proofs (tests/test_fri_oracle.py), which is the reference's own
acceptance criterion for this path (plonky2/src/fri/mod.rs tests run prove -> verify).

Elements of the extension are tuples (c0, c1) of Python ints.  Sizes are kept small (pure-Python loops).
"""
import numpy as np

import oracle
from oracle.quotient import P, inv

W = 7                      # goldilocks_extensions.rs:19
SPONGE_RATE, SPONGE_WIDTH = 8, 12
SALT_SIZE = 4              # fri/oracle.rs:31


# ---- quadratic extension ------------------------------------------------------------------------------
def ext(x):
    return (x % P, 0)


def eadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def esub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def emul(a, b):  # quadratic.rs:176-184
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def escale(a, s):
    return (a[0] * s % P, a[1] * s % P)


def einv(a):  # quadratic.rs:86-99: a^-1 = frobenius(a) / (a * frobenius(a)),  frobenius(a) = (a0, -a1)
    norm = (a[0] * a[0] - W * a[1] * a[1]) % P
    ni = inv(norm)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def epow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = emul(r, a)
        a = emul(a, a)
        e >>= 1
    return r


# ---- Challenger (iop/challenger.rs:15-150) ------------------------------------------------------------
class Challenger:
    def __init__(self, state=None, input_buffer=(), output_buffer=()):
        self.sponge_state = [int(x) for x in state] if state is not None else [0] * SPONGE_WIDTH
        self.input_buffer = [int(x) for x in input_buffer]
        self.output_buffer = [int(x) for x in output_buffer]

    def clone(self):
        return Challenger(self.sponge_state, self.input_buffer, self.output_buffer)

    def observe_element(self, e):
        self.output_buffer = []
        self.input_buffer.append(int(e) % P)
        if len(self.input_buffer) == SPONGE_RATE:
            self.duplexing()

    def observe_elements(self, es):
        for e in es:
            self.observe_element(e)

    def observe_extension_elements(self, es):
        for e in es:
            self.observe_elements(e)

    def observe_cap(self, cap):
        for h in cap:
            self.observe_elements([int(x) for x in h])

    def get_challenge(self):
        if self.input_buffer or not self.output_buffer:
            self.duplexing()
        return self.output_buffer.pop()

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        a = self.get_n_challenges(2)
        return (a[0], a[1])

    def duplexing(self):
        assert len(self.input_buffer) <= SPONGE_RATE
        for i, x in enumerate(self.input_buffer):
            self.sponge_state[i] = x
        self.input_buffer = []
        self.sponge_state = [int(x) for x in oracle.poseidon(np.array(self.sponge_state, dtype=np.uint64))]
        self.output_buffer = list(self.sponge_state[:SPONGE_RATE])


# ---- instance description -----------------------------------------------------------------------------
class FriParams:
    """fri/mod.rs:17-103 (config + degree_bits + reduction_arity_bits)."""

    def __init__(self, degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, reduction_arity_bits, hiding=False):
        self.degree_bits, self.rate_bits, self.cap_height = degree_bits, rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.reduction_arity_bits, self.hiding = list(reduction_arity_bits), hiding

    @property
    def lde_bits(self):
        return self.degree_bits + self.rate_bits

    @property
    def final_poly_len(self):
        return 1 << (self.degree_bits - sum(self.reduction_arity_bits))


def constant_arity_bits(arity_bits, final_poly_bits, degree_bits, rate_bits, cap_height):
    """FriReductionStrategy::ConstantArityBits (fri/reduction_strategies.rs:38-49)."""
    out = []
    while degree_bits > final_poly_bits and degree_bits + rate_bits - arity_bits >= cap_height:
        out.append(arity_bits)
        degree_bits -= arity_bits
    return out


class FriBatchInfo:
    """fri/structure.rs:34-38: a point and the (oracle_index, polynomial_index) pairs opened there."""

    def __init__(self, point, polynomials):
        self.point, self.polynomials = point, list(polynomials)


# ---- openings (plonk/proof.rs:305-334) ----------------------------------------------------------------
def eval_poly_base_at_ext(coeffs, z):
    acc = (0, 0)
    for c in reversed([int(c) for c in coeffs]):
        acc = emul(acc, z)
        acc = ((acc[0] + c) % P, acc[1])
    return acc


def fri_openings(batches, oracles):
    """FriOpenings (fri/structure.rs:66-74): for each batch the value of each of its polynomials at its point."""
    return [[eval_poly_base_at_ext(oracles[o].coeffs[p], b.point) for (o, p) in b.polynomials] for b in batches]


# ---- prove_openings (fri/oracle.rs:1046-1110) ---------------------------------------------------------
def final_poly_coeffs(batches, oracles, alpha):
    """The polynomial that goes into FRI: sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i), times X."""
    n = oracles[0].coeffs.shape[1]
    final = []
    count = 0                                                   # ReducingFactor.count
    for b in batches:
        comp = [(0, 0)] * n                                     # reduce_polys_base (reducing.rs:87-100)
        apow = (1, 0)
        for (o, p) in b.polynomials:
            col = oracles[o].coeffs[p]
            comp = [eadd(c, escale(apow, int(x))) for c, x in zip(comp, col)]
            apow = emul(apow, alpha)
            count += 1
        # divide_by_linear (division.rs:75-88)
        bs, acc = [], (0, 0)
        for c in reversed(comp):
            acc = eadd(emul(acc, b.point), c)
            bs.append(acc)
        bs.pop()
        bs.reverse()
        shift = epow(alpha, count)                              # shift_poly (reducing.rs:108-111)
        count = 0
        final = [emul(c, shift) for c in final]
        if len(final) < len(bs):
            final = final + [(0, 0)] * (len(bs) - len(final))
        final = [eadd(a, q) for a, q in zip(final, bs)]
    return [(0, 0)] + final                                     # coeffs.insert(0, ZERO), oracle.rs:1084


def ext_coset_fft(coeffs, shift):
    """coset_fft of an extension polynomial = two base-field coset FFTs (the twiddles and the shift are in the base field)."""
    c0 = oracle.coset_fft(np.array([c[0] for c in coeffs], dtype=np.uint64), shift)
    c1 = oracle.coset_fft(np.array([c[1] for c in coeffs], dtype=np.uint64), shift)
    return [(int(a), int(b)) for a, b in zip(c0, c1)]


def bitrev_list(v):
    bits = (len(v)).bit_length() - 1
    return [v[oracle.reverse_bits(i, bits)] for i in range(len(v))] if bits else list(v)


class Tree:
    def __init__(self, leaves, cap_height):
        self.leaves = np.array(leaves, dtype=np.uint64)
        self.cap_height = cap_height
        self.digests, self.cap = oracle.merkle_tree(self.leaves, cap_height)

    def prove(self, i):
        return oracle.merkle_prove(self.digests, self.leaves.shape[0], self.cap_height, i)


def fri_committed_trees(coeffs, values, challenger, params):
    """fri/prover.rs:76-120."""
    trees, betas = [], []
    shift = 7
    for arity_bits in params.reduction_arity_bits:
        arity = 1 << arity_bits
        values = bitrev_list(values)
        leaves = [[x for e in values[i:i + arity] for x in e] for i in range(0, len(values), arity)]   # flatten
        tree = Tree(leaves, params.cap_height)
        challenger.observe_cap(tree.cap)
        trees.append(tree)
        beta = challenger.get_extension_challenge()
        betas.append(beta)
        folded = []
        for i in range(0, len(coeffs), arity):                  # reduce_with_powers(chunk, beta)
            acc = (0, 0)
            for c in reversed(coeffs[i:i + arity]):
                acc = eadd(emul(acc, beta), c)
            folded.append(acc)
        coeffs = folded
        shift = pow(shift, arity, P)
        values = ext_coset_fft(coeffs, shift)
    removed = coeffs[len(coeffs) >> params.rate_bits:]
    assert all(c == (0, 0) for c in removed), "the truncated coefficients must be zero (prover.rs:111-115)"
    coeffs = coeffs[:len(coeffs) >> params.rate_bits]
    challenger.observe_extension_elements(coeffs)
    return trees, coeffs, betas


def pow_min_leading_zeros(params):
    return params.proof_of_work_bits + (64 - 64)  # F::order().bits() = 64 (prover.rs:128)


def fri_proof_of_work(challenger, params, start=0):
    """fri/prover.rs:123-171.  The reference takes ANY satisfying candidate (rayon find_any); this restatement and
    the device return the smallest one, which is the deterministic choice among the reference's possible outputs."""
    min_lz = pow_min_leading_zeros(params)
    state = list(challenger.sponge_state)
    pos = len(challenger.input_buffer)
    for i, x in enumerate(challenger.input_buffer):
        state[i] = x
    cand = start
    while True:
        s = list(state)
        s[pos] = cand
        out = oracle.poseidon(np.array(s, dtype=np.uint64))
        resp = int(out[SPONGE_RATE - 1]) % P
        if 64 - resp.bit_length() >= min_lz:
            break
        cand += 1
    challenger.observe_element(cand)
    resp = challenger.get_challenge()
    assert 64 - (resp % P).bit_length() >= min_lz
    return cand


def fri_prover_query_rounds(initial_batches, trees, challenger, n, params):
    """fri/prover.rs:173-260.  Returns (indices, rounds); a round = (initial [(row, siblings)], steps [(evals, siblings)])."""
    challs = challenger.get_n_challenges(params.num_query_rounds)
    indices, rounds = [], []
    for r in challs:
        x_index = r % n
        indices.append(x_index)
        initial = [(np.array(b.leaves[x_index]), oracle.merkle_prove(b.digests, b.leaves.shape[0], b.cap_height, x_index))
                   for b in initial_batches]
        steps = []
        xi = x_index
        for i, tree in enumerate(trees):
            ab = params.reduction_arity_bits[i]
            steps.append((np.array(tree.leaves[xi >> ab]), tree.prove(xi >> ab)))
            xi >>= ab
        rounds.append((initial, steps))
    return indices, rounds


class FriProof:
    """fri/proof.rs: commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness (+ what the tests compare)."""


def prove_openings(batches, oracles, challenger, params):
    """fri/oracle.rs:1046-1110 + fri/prover.rs:23-70.  `oracles`: oracle.Batch objects with .cap_height set."""
    alpha = challenger.get_extension_challenge()
    final = final_poly_coeffs(batches, oracles, alpha)
    n = len(final)
    lde = final + [(0, 0)] * ((n << params.rate_bits) - n)       # lde(rate_bits), polynomial/mod.rs:205-207
    values = ext_coset_fft(lde, 7)
    pr = FriProof()
    pr.alpha = alpha
    pr.final_poly_in = final
    pr.lde_values = values
    trees, final_coeffs, betas = fri_committed_trees(lde, values, challenger, params)
    pr.trees, pr.betas = trees, betas
    pr.commit_phase_merkle_caps = [t.cap for t in trees]
    pr.final_poly = final_coeffs
    pr.pow_witness = fri_proof_of_work(challenger, params)
    pr.query_indices, pr.query_round_proofs = fri_prover_query_rounds(oracles, trees, challenger, len(lde), params)
    return pr


# ---- verifier (fri/verifier.rs) -----------------------------------------------------------------------
def fri_challenges(challenger, caps, final_poly, pow_witness, params):
    """fri/challenges.rs:24-66."""
    alpha = challenger.get_extension_challenge()
    betas = []
    for cap in caps:
        challenger.observe_cap(cap)
        betas.append(challenger.get_extension_challenge())
    challenger.observe_extension_elements(final_poly)
    challenger.observe_element(pow_witness)
    pow_response = challenger.get_challenge()
    lde_size = 1 << params.lde_bits
    indices = [challenger.get_challenge() % lde_size for _ in range(params.num_query_rounds)]
    return alpha, betas, pow_response, indices


def interpolate(points, x):
    """field/src/interpolation.rs:31-51 (the value of the unique interpolant; plain Lagrange)."""
    for xi, yi in points:
        if xi == x:
            return yi
    total = (0, 0)
    for i, (xi, yi) in enumerate(points):
        num, den = (1, 0), (1, 0)
        for j, (xj, _) in enumerate(points):
            if j != i:
                num = emul(num, esub(x, xj))
                den = emul(den, esub(xi, xj))
        total = eadd(total, emul(yi, emul(num, einv(den))))
    return total


def compute_evaluation(x, x_index_within_coset, arity_bits, evals, beta):
    """fri/verifier.rs:20-46."""
    arity = 1 << arity_bits
    g = oracle.primitive_root_of_unity(arity_bits)
    evals = bitrev_list(list(evals))
    rev = oracle.reverse_bits(x_index_within_coset, arity_bits)
    coset_start = x * pow(g, arity - rev, P) % P
    points = [(ext(coset_start * pow(g, i, P)), evals[i]) for i in range(arity)]
    return interpolate(points, beta)


def verify_fri_proof(batches, oracle_salted, openings, challenges, initial_caps, proof, params):
    """fri/verifier.rs:62-241.  Raises AssertionError like the reference's `ensure!`.  `oracle_salted[i]`: the i-th
    oracle's leaves carry SALT_SIZE blinding columns (instance.oracles[i].blinding && params.hiding)."""
    alpha, betas, pow_response, indices = challenges
    n_log = params.lde_bits
    assert 64 - (pow_response % P).bit_length() >= pow_min_leading_zeros(params), "Invalid proof of work witness."
    assert len(proof.query_round_proofs) == params.num_query_rounds
    assert len(proof.final_poly) == params.final_poly_len
    # PrecomputedReducedOpenings::from_os_and_alpha (verifier.rs:250-262)
    reduced_openings = []
    for vals in openings:
        acc = (0, 0)
        for v in reversed(vals):
            acc = eadd(emul(acc, alpha), v)
        reduced_openings.append(acc)
    for x_index, (initial, steps) in zip(indices, proof.query_round_proofs):
        for (row, sib), cap in zip(initial, initial_caps):                      # fri_verify_initial_proof
            assert oracle.merkle_verify(row, x_index, cap, sib), "Invalid Merkle proof."
        subgroup_x = 7 * pow(oracle.primitive_root_of_unity(n_log), oracle.reverse_bits(x_index, n_log), P) % P
        # fri_combine_initial (verifier.rs:117-163)
        total, count = (0, 0), 0
        for b, red in zip(batches, reduced_openings):
            evals = []
            for (o, p) in b.polynomials:
                row = initial[o][0]
                if oracle_salted[o]:
                    row = row[:len(row) - SALT_SIZE]
                evals.append(ext(int(row[p])))
            acc = (0, 0)
            for v in reversed(evals):
                acc = eadd(emul(acc, alpha), v)
                count += 1
            numerator = esub(acc, red)
            denominator = esub(ext(subgroup_x), b.point)
            total = emul(total, epow(alpha, count))
            count = 0
            total = eadd(total, emul(numerator, einv(denominator)))
        old_eval = emul(total, ext(subgroup_x))
        xi = x_index
        for i, ab in enumerate(params.reduction_arity_bits):
            arity = 1 << ab
            flat, sib = steps[i]
            evals = [(int(flat[2 * j]), int(flat[2 * j + 1])) for j in range(arity)]
            coset_index, within = xi >> ab, xi & (arity - 1)
            assert evals[within] == old_eval, "FRI fold consistency"
            old_eval = compute_evaluation(subgroup_x, within, ab, evals, betas[i])
            assert oracle.merkle_verify(np.array(flat, dtype=np.uint64), coset_index, proof.commit_phase_merkle_caps[i], sib), \
                "Invalid Merkle proof."
            subgroup_x = pow(subgroup_x, arity, P)
            xi = coset_index
        acc = (0, 0)
        for c in reversed(proof.final_poly):
            acc = eadd(emul(acc, ext(subgroup_x)), c)
        assert acc == old_eval, "Final polynomial evaluation is invalid."
    return True
