"""Synthetic circuits + honest witnesses for the quotient tests (shared by the CPU oracle tests and the GPU parity tests).

The witness rows restate what each gate's generator writes (e.g. arithmetic_u32.rs:420-470, comparison.rs:404-520,
random_access.rs:453-520, base_sum.rs:233-280), the copy constraints are the identity permutation (sigma_j(x) = k_j x,
so Z = 1 and every partial product is 1), and the selector polynomials follow gates/selectors.rs:37-112.
"""
import numpy as np

import oracle
from oracle import quotient as Q
from oracle import recursion_gates as RG

P = Q.P


def rnd(rng, bound=P):
    return int(rng.integers(0, bound, dtype=np.uint64)) if bound > 2**62 else int(rng.integers(0, bound))


def honest_row(gate, rng, num_wires, consts, pih):
    """A row satisfying all of `gate`'s constraints; unused wires random.  `consts`: the gate's constants (writable)."""
    w = [rnd(rng) for _ in range(num_wires)]
    if isinstance(gate, Q.NoopGate):
        pass
    elif isinstance(gate, Q.ConstantGate):
        for i in range(gate.num_consts):
            w[i] = consts[i]
    elif isinstance(gate, Q.PublicInputGate):
        w[:4] = pih
    elif isinstance(gate, Q.ArithmeticGate):
        for i in range(gate.num_ops):
            m0, m1, add = w[4 * i], w[4 * i + 1], w[4 * i + 2]
            w[4 * i + 3] = (m0 * m1 * consts[0] + add * consts[1]) % P
    elif isinstance(gate, Q.BaseSumGate):
        limbs = [rnd(rng, gate.base) for _ in range(gate.num_limbs)]
        w[1:1 + gate.num_limbs] = limbs
        w[0] = sum(l * gate.base**i for i, l in enumerate(limbs)) % P
    elif isinstance(gate, Q.PoseidonGate):
        w = gate.honest_row([rnd(rng) for _ in range(12)], rnd(rng, 2), num_wires)
        for i in range(Q.PoseidonGate.START_FULL_1 + 48, num_wires):
            w[i] = rnd(rng)
    elif isinstance(gate, Q.RandomAccessGate):
        vs = gate.vec_size()
        for copy in range(gate.num_copies):
            base = (2 + vs) * copy
            idx = rnd(rng, vs)
            w[base] = idx
            w[base + 1] = w[base + 2 + idx]
            for i in range(gate.bits):
                w[gate.num_routed() + copy * gate.bits + i] = (idx >> i) & 1
        for i in range(gate.num_extra_constants):
            w[(2 + vs) * gate.num_copies + i] = consts[i]
    elif isinstance(gate, Q.U32ArithmeticGate):
        for i in range(gate.num_ops):
            m0, m1, add = rnd(rng, 2**32), rnd(rng, 2**32), rnd(rng, 2**32)
            if i == 0:
                m0 = m1 = add = 2**32 - 1  # hi = 2^32 - 2, exercises the inverse wire near the edge
            routed, limbs = gate.honest_op(m0, m1, add)
            w[6 * i:6 * i + 6] = routed
            w[6 * gate.num_ops + 32 * i:6 * gate.num_ops + 32 * i + 32] = limbs
    elif isinstance(gate, Q.U32AddManyGate):
        na = gate.num_addends
        for i in range(gate.num_ops):
            b = (na + 3) * i
            adds = [rnd(rng, 2**32) for _ in range(na)]
            carry = rnd(rng, 2**32)
            tot = sum(adds) + carry
            w[b:b + na] = adds
            w[b + na], w[b + na + 1], w[b + na + 2] = carry, tot & 0xFFFFFFFF, tot >> 32
            for j in range(18):
                w[(na + 3) * gate.num_ops + 18 * i + j] = (tot >> (2 * j)) & 3
    elif isinstance(gate, Q.U32RangeCheckGate):
        n = gate.num_input_limbs
        for i in range(n):
            v = rnd(rng, 2**32)
            w[i] = v
            for j in range(16):
                w[n + 16 * i + j] = (v >> (2 * j)) & 3
    elif isinstance(gate, Q.U32SubtractionGate):
        for i in range(gate.num_ops):
            x, y, br = rnd(rng, 2**32), rnd(rng, 2**32), rnd(rng, 2)
            d = x - y - br
            ob = 1 if d < 0 else 0
            res = d + ob * 2**32
            w[5 * i:5 * i + 5] = [x, y, br, res, ob]
            for j in range(16):
                w[5 * gate.num_ops + 16 * i + j] = (res >> (2 * j)) & 3
    elif isinstance(gate, Q.ComparisonGate):
        nc, cb = gate.num_chunks, gate.chunk_bits()
        a, b = rnd(rng, 2**gate.num_bits), rnd(rng, 2**gate.num_bits)
        if rnd(rng, 4) == 0:
            b = a
        fc = [(a >> (cb * i)) & ((1 << cb) - 1) for i in range(nc)]
        sc = [(b >> (cb * i)) & ((1 << cb) - 1) for i in range(nc)]
        w[0], w[1] = a, b
        w[4:4 + nc], w[4 + nc:4 + 2 * nc] = fc, sc
        msd = 0
        for i in range(nc):
            diff = (sc[i] - fc[i]) % P
            eq = 1 if diff == 0 else 0
            w[4 + 2 * nc + i] = Q.inv(diff) if diff else 1     # equality dummy
            w[4 + 3 * nc + i] = eq
            w[4 + 4 * nc + i] = eq * msd % P                   # intermediate value
            msd = (w[4 + 4 * nc + i] + (1 - eq) * diff) % P
        w[3] = msd
        # msd + 2^cb as cb+1 bits (msd is a small signed value: second - first chunk in (-2^cb, 2^cb))
        val = (msd + (1 << cb)) % P
        bits = [(val >> i) & 1 for i in range(cb + 1)]
        w[4 + 5 * nc:4 + 5 * nc + cb + 1] = bits
        w[2] = bits[cb]
    elif isinstance(gate, RECURSION_GATES):
        w = gate.honest_row(lambda: rnd(rng), num_wires, consts)
    else:
        raise TypeError(gate)
    return w


RECURSION_GATES = (RG.ArithmeticExtensionGate, RG.MulExtensionGate, RG.ReducingGate, RG.ExponentiationGate, RG.PoseidonMdsGate,
                   RG.HighDegreeInterpolationGate, RG.LowDegreeInterpolationGate)


def gate_constants_needed(gate):
    if isinstance(gate, RG.ArithmeticExtensionGate):
        return 2
    if isinstance(gate, RG.MulExtensionGate):
        return 1
    if isinstance(gate, Q.ConstantGate):
        return gate.num_consts
    if isinstance(gate, Q.ArithmeticGate):
        return 2
    if isinstance(gate, Q.RandomAccessGate):
        return gate.num_extra_constants
    return 0


class Instance:
    """A circuit + honest witness + the three committed matrices (values on H, column-major)."""


def build_instance(gates, groups, selector_indices, degree_bits, num_wires, num_routed_wires, seed=0, rate_bits=3,
                   num_challenges=2, quotient_degree_factor=8, corrupt=False):
    rng = np.random.default_rng(seed)
    n = 1 << degree_bits
    num_selectors = len(groups)
    num_gate_consts = max([gate_constants_needed(g) for g in gates] + [2])
    num_constants = num_selectors + num_gate_consts
    k_is = [pow(7, j, P) for j in range(num_routed_wires)]  # distinct coset representatives (powers of the generator)
    circ = Q.Circuit(gates, selector_indices, groups, num_wires, num_routed_wires, num_constants, k_is, degree_bits,
                     rate_bits, num_challenges, quotient_degree_factor)
    pih = [rnd(rng) for _ in range(4)]
    wires = np.zeros((num_wires, n), dtype=np.uint64)
    consts = np.zeros((num_constants, n), dtype=np.uint64)
    for row in range(n):
        gi = row % len(gates) if row < 2 * len(gates) else rnd(rng, len(gates))
        gate = gates[gi]
        gc = [rnd(rng) for _ in range(num_gate_consts)]
        w = honest_row(gate, rng, num_wires, gc, pih)
        wires[:, row] = np.array(w, dtype=np.uint64)
        for s, grp in enumerate(groups):
            consts[s, row] = gi if grp[0] <= gi < grp[1] else Q.UNUSED_SELECTOR  # selectors.rs:89-110
        consts[num_selectors:, row] = np.array(gc, dtype=np.uint64)
    if corrupt:
        wires[0, 1] ^= np.uint64(1)  # row 1 is a ConstantGate row: wire 0 must equal its constant
    wn = oracle.primitive_root_of_unity(degree_bits)
    sigmas = np.zeros((num_routed_wires, n), dtype=np.uint64)
    x = 1
    for row in range(n):
        for j in range(num_routed_wires):
            sigmas[j, row] = k_is[j] * x % P
        x = x * wn % P
    npp = circ.num_partial_products
    zs_pp = np.ones((num_challenges * (1 + npp), n), dtype=np.uint64)
    inst = Instance()
    inst.circ, inst.pih = circ, pih
    inst.wires, inst.consts_sigmas, inst.zs_pp = wires, np.concatenate([consts, sigmas]), zs_pp
    inst.betas = [rnd(rng) for _ in range(num_challenges)]
    inst.gammas = [rnd(rng) for _ in range(num_challenges)]
    inst.alphas = [rnd(rng) for _ in range(num_challenges)]
    return inst


def standard_gate_sets(num_wires=135, num_routed=80):
    """Two gate mixes: (A) everything but Poseidon in groups small enough to leave degree slack, (B) with Poseidon."""
    mix = [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate(), Q.ArithmeticGate(num_routed // 4),
           Q.BaseSumGate(min(63, num_routed - 1), 2), Q.RandomAccessGate.new_from_config(num_wires, num_routed, 2, 2),
           Q.U32ArithmeticGate(Q.U32ArithmeticGate.num_ops_for(num_wires, num_routed)),
           Q.U32AddManyGate(3, Q.U32AddManyGate.num_ops_for(3, num_wires, num_routed)),
           Q.U32RangeCheckGate(min(7, num_wires // 17)), Q.U32SubtractionGate(Q.U32SubtractionGate.num_ops_for(num_wires, num_routed)),
           Q.ComparisonGate(8, 4)]
    groups_a = [(0, 4), (4, 8), (8, 11)]
    sel_a = [0] * 4 + [1] * 4 + [2] * 3
    with_pos = mix + [Q.PoseidonGate()]
    groups_b = groups_a + [(11, 12)]
    sel_b = sel_a + [3]
    return (mix, groups_a, sel_a), (with_pos, groups_b, sel_b)


def recursion_gate_set(num_wires=135, num_routed=80):
    """The gate types of a recursive-verifier circuit under standard_recursion_config (circuit_builder.rs + the gates
    fri/recursive_verifier.rs and the hashing gadgets add), in groups that keep filter + gate degree within 8."""
    gates = [Q.NoopGate(), Q.ConstantGate(2), Q.PublicInputGate(), Q.ArithmeticGate(num_routed // 4),
             RG.ArithmeticExtensionGate(num_routed // 8), RG.MulExtensionGate(num_routed // 6),
             RG.ReducingGate(min(num_routed - 6, (num_wires - 4) // 3)), RG.ReducingExtensionGate(min((num_routed - 6) // 2, (num_wires - 4) // 4)),
             Q.BaseSumGate(min(63, num_routed - 1), 2), Q.RandomAccessGate.new_from_config(num_wires, num_routed, 4, 2),
             RG.ExponentiationGate(min(num_routed - 2, (num_wires - 2) // 2)), RG.PoseidonMdsGate(),
             RG.LowDegreeInterpolationGate(4), RG.HighDegreeInterpolationGate(2), Q.PoseidonGate()]
    groups = [(0, 5), (5, 9), (9, 12), (12, 14), (14, 15)]
    sel = [0] * 5 + [1] * 4 + [2] * 3 + [3] * 2 + [4]
    return gates, groups, sel
