"""oracle/recursion_gates.py -- TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of the constraint polynomials of the
gates a recursive-verifier circuit adds to the base set, evaluated at BASE-field points
like compute_quotient_polys does (eval_unfiltered_base_one of each gate):

    ArithmeticExtensionGate        plonky2/src/gates/arithmetic_extension.rs:129-147   (wires :40-51)
    MulExtensionGate               plonky2/src/gates/multiplication_extension.rs:122-137 (wires :40-48)
    ReducingGate                   plonky2/src/gates/reducing.rs:160-180               (wires :33-58)
    ReducingExtensionGate          plonky2/src/gates/reducing_extension.rs:160-179     (wires :34-59)
    ExponentiationGate             plonky2/src/gates/exponentiation.rs:266-299         (wires :55-73)
    PoseidonMdsGate                plonky2/src/gates/poseidon_mds.rs:184-203, hash/poseidon.rs mds_layer_field
    HighDegreeInterpolationGate    plonky2/src/gates/high_degree_interpolation.rs:126-147, gates/interpolation.rs:21-76
    LowDegreeInterpolationGate     plonky2/src/gates/low_degree_interpolation.rs:356-404 (wires :50-72)

Extension elements occupy D = 2 consecutive wires (c0, c1) of F[X]/(X^2 - 7).  Each class also produces an honest row
(what the gate's generator writes) so that the tests can restate gate_testing.rs: constraints vanish on honest rows and
not on corrupted ones.  Parity pinning: no stored vectors in the reference; see oracle/quotient.py's header.
"""
from . import poseidon_params as PP
from .quotient import Gate, P, inv

D = 2
W7 = 7


def xadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def xsub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def xmul(a, b):
    return ((a[0] * b[0] + W7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def xscale(a, s):
    return (a[0] * s % P, a[1] * s % P)


def root_of_unity(k):
    b = 1753635133440165772
    for _ in range(32 - k):
        b = b * b % P
    return b


def ext_at(wires, k):
    return (wires[k], wires[k + 1])


class ArithmeticExtensionGate(Gate):
    type_id = 12

    def __init__(self, num_ops):
        self.num_ops, self.params = num_ops, (num_ops,)

    @staticmethod
    def num_ops_for(num_routed):           # new_from_config: num_routed_wires / (4 * D)
        return num_routed // (4 * D)

    def num_constraints(self):
        return self.num_ops * D

    def eval_unfiltered(self, consts, wires, pih):
        c0, c1 = consts[0], consts[1]
        out = []
        for i in range(self.num_ops):
            m0, m1, add, o = (ext_at(wires, 8 * i + 2 * t) for t in range(4))
            out += list(xsub(o, xadd(xscale(xmul(m0, m1), c0), xscale(add, c1))))
        return out

    def honest_row(self, rnd, num_wires, consts):
        w = [rnd() for _ in range(num_wires)]
        for i in range(self.num_ops):
            m0, m1, add = (ext_at(w, 8 * i + 2 * t) for t in range(3))
            w[8 * i + 6], w[8 * i + 7] = xadd(xscale(xmul(m0, m1), consts[0]), xscale(add, consts[1]))
        return w


class MulExtensionGate(Gate):
    type_id = 13

    def __init__(self, num_ops):
        self.num_ops, self.params = num_ops, (num_ops,)

    def num_constraints(self):
        return self.num_ops * D

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        for i in range(self.num_ops):
            m0, m1, o = (ext_at(wires, 6 * i + 2 * t) for t in range(3))
            out += list(xsub(o, xscale(xmul(m0, m1), consts[0])))
        return out

    def honest_row(self, rnd, num_wires, consts):
        w = [rnd() for _ in range(num_wires)]
        for i in range(self.num_ops):
            w[6 * i + 4], w[6 * i + 5] = xscale(xmul(ext_at(w, 6 * i), ext_at(w, 6 * i + 2)), consts[0])
        return w


class ReducingGate(Gate):
    """acc_i = acc_{i-1} * alpha + coeff_i with BASE-field coefficients; the last accumulator is the output."""
    type_id = 14

    def __init__(self, num_coeffs):
        self.num_coeffs, self.params = num_coeffs, (num_coeffs,)

    def num_constraints(self):
        return D * self.num_coeffs

    def coeff(self, wires, i):
        return (wires[3 * D + i], 0)

    def acc_wire(self, i):
        return 0 if i == self.num_coeffs - 1 else 3 * D + self.num_coeffs + D * i

    def eval_unfiltered(self, consts, wires, pih):
        alpha, acc = ext_at(wires, D), ext_at(wires, 2 * D)
        out = []
        for i in range(self.num_coeffs):
            a = ext_at(wires, self.acc_wire(i))
            out += list(xsub(xadd(xmul(acc, alpha), self.coeff(wires, i)), a))
            acc = a
        return out

    def min_wires(self):
        return 3 * D + self.num_coeffs + D * (self.num_coeffs - 1)

    def honest_row(self, rnd, num_wires, consts):
        w = [rnd() for _ in range(num_wires)]
        alpha, acc = ext_at(w, D), ext_at(w, 2 * D)
        for i in range(self.num_coeffs):
            acc = xadd(xmul(acc, alpha), self.coeff(w, i))
            k = self.acc_wire(i)
            w[k], w[k + 1] = acc
        return w


class ReducingExtensionGate(ReducingGate):
    """Same with extension-field coefficients (D wires each)."""
    type_id = 15

    def coeff(self, wires, i):
        return ext_at(wires, 3 * D + D * i)

    def acc_wire(self, i):
        return 0 if i == self.num_coeffs - 1 else 3 * D + D * self.num_coeffs + D * i

    def min_wires(self):
        return 3 * D + D * self.num_coeffs + D * (self.num_coeffs - 1)


class ExponentiationGate(Gate):
    """output = base^(sum bits_i 2^i): square-and-multiply from the most significant bit, one intermediate per bit."""
    type_id = 16

    def __init__(self, num_power_bits):
        self.num_power_bits, self.params = num_power_bits, (num_power_bits,)

    def num_constraints(self):
        return self.num_power_bits + 1

    def eval_unfiltered(self, consts, wires, pih):
        nb = self.num_power_bits
        base, output = wires[0], wires[1 + nb]
        inter = wires[2 + nb:2 + 2 * nb]
        out = []
        for i in range(nb):
            prev = 1 if i == 0 else inter[i - 1] * inter[i - 1] % P
            bit = wires[1 + (nb - 1 - i)]
            out.append((prev * (bit * base + (1 - bit)) - inter[i]) % P)
        out.append((output - inter[nb - 1]) % P)
        return out

    def honest_row(self, rnd, num_wires, consts):
        nb = self.num_power_bits
        w = [rnd() for _ in range(num_wires)]
        bits = [rnd() & 1 for _ in range(nb)]
        w[1:1 + nb] = bits
        cur = 1
        for i in range(nb):
            prev = 1 if i == 0 else cur * cur % P
            cur = prev * (w[0] if bits[nb - 1 - i] else 1) % P
            w[2 + nb + i] = cur
        w[1 + nb] = cur
        return w


class PoseidonMdsGate(Gate):
    """outputs = MDS * inputs on 12 extension elements (the MDS matrix is over the base field: it acts on c0 and c1 alike)."""
    type_id = 17

    def num_constraints(self):
        return 12 * D

    @staticmethod
    def mds(v):
        out = []
        for r in range(12):
            acc = (0, 0)
            for i in range(12):
                acc = xadd(acc, xscale(v[(i + r) % 12], PP.MDS_CIRC[i]))
            out.append(xadd(acc, xscale(v[r], PP.MDS_DIAG[r])))
        return out

    def eval_unfiltered(self, consts, wires, pih):
        computed = self.mds([ext_at(wires, D * i) for i in range(12)])
        out = []
        for i in range(12):
            out += list(xsub(ext_at(wires, D * (12 + i)), computed[i]))
        return out

    def honest_row(self, rnd, num_wires, consts):
        w = [rnd() for _ in range(num_wires)]
        for i, c in enumerate(self.mds([ext_at(w, D * i) for i in range(12)])):
            w[D * (12 + i)], w[D * (12 + i) + 1] = c
        return w


class _InterpolationLayout:
    """gates/interpolation.rs:21-76"""

    def __init__(self, subgroup_bits):
        self.subgroup_bits, self.params = subgroup_bits, (subgroup_bits,)
        self.np = 1 << subgroup_bits
        self.start_values = 1
        self.start_eval_point = 1 + self.np * D
        self.start_eval_value = self.start_eval_point + D
        self.start_coeffs = self.start_eval_value + D
        self.end_coeffs = self.start_coeffs + D * self.np

    def value(self, wires, i):
        return ext_at(wires, self.start_values + D * i)

    def coeff(self, wires, i):
        return ext_at(wires, self.start_coeffs + D * i)

    def interpolate_row(self, rnd, num_wires):
        """A row whose coefficients interpolate random values on the coset shift*<g> (what InterpolationGenerator writes)."""
        w = [rnd() for _ in range(num_wires)]
        coeffs = [(rnd(), rnd()) for _ in range(self.np)]
        shift, g = w[0], root_of_unity(self.subgroup_bits)
        for i in range(self.np):
            k = self.start_coeffs + D * i
            w[k], w[k + 1] = coeffs[i]
        for i in range(self.np):
            pt = shift * pow(g, i, P) % P
            acc = (0, 0)
            for c in reversed(coeffs):
                acc = xadd(xscale(acc, pt), c)
            k = self.start_values + D * i
            w[k], w[k + 1] = acc
        ep = ext_at(w, self.start_eval_point)
        acc = (0, 0)
        for c in reversed(coeffs):
            acc = xadd(xmul(acc, ep), c)
        w[self.start_eval_value], w[self.start_eval_value + 1] = acc
        return w, coeffs, ep


class HighDegreeInterpolationGate(_InterpolationLayout, Gate):
    type_id = 18

    def num_constraints(self):
        return self.np * D + D

    def min_wires(self):
        return self.end_coeffs

    def eval_unfiltered(self, consts, wires, pih):
        coeffs = [self.coeff(wires, i) for i in range(self.np)]
        shift, g = wires[0], root_of_unity(self.subgroup_bits)
        out = []
        for i in range(self.np):
            pt = shift * pow(g, i, P) % P
            acc = (0, 0)
            for c in reversed(coeffs):                    # eval_base: acc.scalar_mul(x) + c
                acc = xadd(xscale(acc, pt), c)
            out += list(xsub(self.value(wires, i), acc))
        ep = ext_at(wires, self.start_eval_point)
        acc = (0, 0)
        for c in reversed(coeffs):
            acc = xadd(xmul(acc, ep), c)
        out += list(xsub(ext_at(wires, self.start_eval_value), acc))
        return out

    def honest_row(self, rnd, num_wires, consts):
        return self.interpolate_row(rnd, num_wires)[0]


class LowDegreeInterpolationGate(_InterpolationLayout, Gate):
    """All constraints of degree <= 2: the powers of the shift and of the evaluation point are witnessed."""
    type_id = 19

    def num_constraints(self):
        return (self.np - 2) + self.np * D + (self.np - 2) * D + D

    def shift_power_wire(self, i):        # low_degree_interpolation.rs:50-57
        return 0 if i == 1 else self.end_coeffs + i - 2

    def eval_power_wire(self, i):         # :59-67
        return self.start_eval_point if i == 1 else self.end_coeffs + self.np - 2 + (i - 2) * D

    def min_wires(self):
        return self.eval_power_wire(self.np - 1) + D

    def eval_unfiltered(self, consts, wires, pih):
        np_ = self.np
        coeffs = [self.coeff(wires, i) for i in range(np_)]
        ps = [wires[self.shift_power_wire(i)] for i in range(1, np_)]
        shift = ps[0]
        out = [(ps[i - 1] * shift - ps[i]) % P for i in range(1, np_ - 1)]
        ps = [1] + ps
        altered = [xscale(c, p) for c, p in zip(coeffs, ps)]
        g = root_of_unity(self.subgroup_bits)
        for i in range(np_):
            pt = pow(g, i, P)
            acc = (0, 0)
            for c in reversed(altered):
                acc = xadd(xscale(acc, pt), c)
            out += list(xsub(self.value(wires, i), acc))
        epp = [ext_at(wires, self.eval_power_wire(i)) for i in range(1, np_)]
        ep = epp[0]
        for i in range(1, np_ - 1):
            out += list(xsub(xmul(epp[i - 1], ep), epp[i]))
        acc = coeffs[0]                                   # eval_with_powers (polynomial/mod.rs:169-176)
        for c, pw in zip(coeffs[1:], epp):
            acc = xadd(acc, xmul(pw, c))
        out += list(xsub(ext_at(wires, self.start_eval_value), acc))
        return out

    def honest_row(self, rnd, num_wires, consts):
        w, coeffs, ep = self.interpolate_row(rnd, num_wires)
        cur, curx = w[0], ep
        for i in range(2, self.np):
            cur = cur * w[0] % P
            w[self.shift_power_wire(i)] = cur
            curx = xmul(curx, ep)
            k = self.eval_power_wire(i)
            w[k], w[k + 1] = curx
        return w
