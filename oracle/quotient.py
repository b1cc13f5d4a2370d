"""oracle/quotient.py -- TEST INFRASTRUCTURE ONLY.  Pure-Python (exact integer) restatement of the reference's
quotient-polynomial evaluation over the LDE domain:

    compute_quotient_polys                  plonky2/src/plonk/prover.rs:790-1034
    eval_vanishing_poly_base_batch          plonky2/src/plonk/vanishing_poly.rs:100-226
    evaluate_gate_constraints_base_batch    plonky2/src/plonk/vanishing_poly.rs:267-306
    Gate::eval_filtered_base_batch, filter  plonky2/src/gates/gate.rs:113-150, 261-268
    check_partial_products                  plonky2/src/util/partial_products.rs:52-78
    reduce_with_powers_multi                plonky2/src/plonk/plonk_common.rs:97-114
    ZeroPolyOnCoset                         field/src/zero_poly_coset.rs:7-60
and of the gates' constraint polynomials (each class cites its file).

Written from the Rust sources, independently of the CUDA kernels (different language, no shared code), for small
circuits.  Parity pinning: the reference stores no vectors for quotient values; this restatement is pinned by the
reference ITSELF run on the GPU box (oracle/_ref = the reference's CUDA translation unit compiled unmodified):
tests/test_ref_cuda_crosscheck.py checks every base gate type against the reference's device-side evaluators
(cuda/*Gate.cuh) and the full quotient values against the reference's compute_quotient_values_kernel on its own 25-gate
circuit.  In addition tests/test_quotient_oracle.py restates the reference's property tests -- every gate's constraints
vanish on a witness row produced by that gate's generator logic and do not on a corrupted one (gate_testing.rs), and the
quotient of an honest witness has its top coefficients zero.
"""
from . import poseidon_params as PP

P = 0xFFFFFFFF00000001
UNUSED_SELECTOR = 0xFFFFFFFF  # gates/selectors.rs:11


def inv(a):
    return pow(a % P, P - 2, P)


def reduce_with_powers(terms, alpha):
    """plonk_common.rs:116-128: sum = sum * alpha + term, from the last term"""
    s = 0
    for t in reversed(terms):
        s = (s * alpha + t) % P
    return s


# ---------------------------------------------------------------------------------------------------------------
# gates
# ---------------------------------------------------------------------------------------------------------------
class Gate:
    type_id = -1
    params = ()

    def num_constraints(self):
        raise NotImplementedError

    def eval_unfiltered(self, consts, wires, pih):
        """consts: local constants with the selector prefix removed; wires: the row; pih: public input hash (4)."""
        raise NotImplementedError


class NoopGate(Gate):
    """gates/noop.rs: no constraints"""
    type_id = 0

    def num_constraints(self):
        return 0

    def eval_unfiltered(self, consts, wires, pih):
        return []


class ConstantGate(Gate):
    """gates/constant.rs:150-158"""
    type_id = 1

    def __init__(self, num_consts):
        self.num_consts, self.params = num_consts, (num_consts,)

    def num_constraints(self):
        return self.num_consts

    def eval_unfiltered(self, consts, wires, pih):
        return [(consts[i] - wires[i]) % P for i in range(self.num_consts)]


class PublicInputGate(Gate):
    """gates/public_input.rs:129-139: wires 0..4 equal the public-inputs hash"""
    type_id = 2

    def num_constraints(self):
        return 4

    def eval_unfiltered(self, consts, wires, pih):
        return [(wires[i] - pih[i]) % P for i in range(4)]


class ArithmeticGate(Gate):
    """gates/arithmetic_base.rs:199-220: out = c0 * x * y + c1 * z, wires 4i..4i+3"""
    type_id = 3

    def __init__(self, num_ops):
        self.num_ops, self.params = num_ops, (num_ops,)

    def num_constraints(self):
        return self.num_ops

    def eval_unfiltered(self, consts, wires, pih):
        c0, c1 = consts[0], consts[1]
        out = []
        for i in range(self.num_ops):
            m0, m1, add, o = wires[4 * i:4 * i + 4]
            out.append((o - (m0 * m1 * c0 + add * c1)) % P)
        return out


class BaseSumGate(Gate):
    """gates/base_sum.rs:213-230 (base B): wire 0 = sum, limbs at 1..1+num_limbs"""
    type_id = 4

    def __init__(self, num_limbs, base=2):
        self.num_limbs, self.base, self.params = num_limbs, base, (num_limbs, base)

    def num_constraints(self):
        return 1 + self.num_limbs

    def eval_unfiltered(self, consts, wires, pih):
        limbs = wires[1:1 + self.num_limbs]
        out = [(reduce_with_powers(limbs, self.base) - wires[0]) % P]
        for limb in limbs:
            prod = 1
            for i in range(self.base):
                prod = prod * (limb - i) % P
            out.append(prod)
        return out


class PoseidonGate(Gate):
    """gates/poseidon.rs:485-564"""
    type_id = 5
    W = 12
    WIRE_SWAP = 24
    START_DELTA = 25
    START_FULL_0 = 29
    START_PARTIAL = 29 + 12 * 3
    START_FULL_1 = 29 + 36 + 22

    def num_constraints(self):
        return 1 + 4 + 12 * 3 + 22 + 12 * 4 + 12  # = 123 (poseidon.rs: num_constraints)

    @staticmethod
    def _sbox(x):
        return pow(x, 7, P)

    @staticmethod
    def _mds(state):
        return [(sum(state[(i + r) % 12] * PP.MDS_CIRC[i] for i in range(12)) + state[r] * PP.MDS_DIAG[r]) % P for r in range(12)]

    @staticmethod
    def _partial_init(state):
        res = [0] * 12
        res[0] = state[0]
        for r in range(1, 12):
            for c in range(1, 12):
                res[c] = (res[c] + state[r] * PP.PARTIAL_INIT_MATRIX[(r - 1) * 11 + (c - 1)]) % P
        return res

    @staticmethod
    def _partial_fast(state, r):
        d = state[0] * (PP.MDS_CIRC[0] + PP.MDS_DIAG[0])
        for i in range(1, 12):
            d += state[i] * PP.PARTIAL_W_HATS[r * 11 + i - 1]
        res = [d % P]
        for i in range(1, 12):
            res.append((state[i] + state[0] * PP.PARTIAL_VS[r * 11 + i - 1]) % P)
        return res

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        swap = wires[self.WIRE_SWAP]
        out.append(swap * (swap - 1) % P)
        for i in range(4):
            out.append((swap * (wires[i + 4] - wires[i]) - wires[self.START_DELTA + i]) % P)
        state = [0] * 12
        for i in range(4):
            d = wires[self.START_DELTA + i]
            state[i] = (wires[i] + d) % P
            state[i + 4] = (wires[i + 4] - d) % P
        for i in range(8, 12):
            state[i] = wires[i]
        rc = 0
        for r in range(4):
            state = [(state[i] + PP.ROUND_CONSTANTS[i + 12 * rc]) % P for i in range(12)]
            if r != 0:
                for i in range(12):
                    sbox_in = wires[self.START_FULL_0 + 12 * (r - 1) + i]
                    out.append((state[i] - sbox_in) % P)
                    state[i] = sbox_in
            state = self._mds([self._sbox(x) for x in state])
            rc += 1
        state = [(state[i] + PP.PARTIAL_FIRST_RC[i]) % P for i in range(12)]
        state = self._partial_init(state)
        for r in range(21):
            sbox_in = wires[self.START_PARTIAL + r]
            out.append((state[0] - sbox_in) % P)
            state[0] = (self._sbox(sbox_in) + PP.PARTIAL_RC[r]) % P
            state = self._partial_fast(state, r)
        sbox_in = wires[self.START_PARTIAL + 21]
        out.append((state[0] - sbox_in) % P)
        state[0] = self._sbox(sbox_in)
        state = self._partial_fast(state, 21)
        rc += 22
        for r in range(4):
            state = [(state[i] + PP.ROUND_CONSTANTS[i + 12 * rc]) % P for i in range(12)]
            for i in range(12):
                sbox_in = wires[self.START_FULL_1 + 12 * r + i]
                out.append((state[i] - sbox_in) % P)
                state[i] = sbox_in
            state = self._mds([self._sbox(x) for x in state])
            rc += 1
        for i in range(12):
            out.append((state[i] - wires[12 + i]) % P)
        return out

    def honest_row(self, inputs, swap, num_wires):
        """The witness PoseidonGenerator produces (gates/poseidon.rs:638-730): used by the tests."""
        w = [0] * num_wires
        w[:12] = [x % P for x in inputs]
        w[self.WIRE_SWAP] = swap
        for i in range(4):
            w[self.START_DELTA + i] = swap * (w[i + 4] - w[i]) % P
        state = list(w[:12])
        if swap:
            state[:4], state[4:8] = state[4:8], state[:4]
        rc = 0
        for r in range(4):
            state = [(state[i] + PP.ROUND_CONSTANTS[i + 12 * rc]) % P for i in range(12)]
            if r != 0:
                for i in range(12):
                    w[self.START_FULL_0 + 12 * (r - 1) + i] = state[i]
            state = self._mds([self._sbox(x) for x in state])
            rc += 1
        state = [(state[i] + PP.PARTIAL_FIRST_RC[i]) % P for i in range(12)]
        state = self._partial_init(state)
        for r in range(22):
            w[self.START_PARTIAL + r] = state[0]
            state[0] = self._sbox(state[0])
            if r < 21:
                state[0] = (state[0] + PP.PARTIAL_RC[r]) % P
            state = self._partial_fast(state, r)
        rc += 22
        for r in range(4):
            state = [(state[i] + PP.ROUND_CONSTANTS[i + 12 * rc]) % P for i in range(12)]
            for i in range(12):
                w[self.START_FULL_1 + 12 * r + i] = state[i]
            state = self._mds([self._sbox(x) for x in state])
            rc += 1
        w[12:24] = state
        return w


class RandomAccessGate(Gate):
    """gates/random_access.rs:409-450"""
    type_id = 6

    def __init__(self, bits, num_copies, num_extra_constants):
        self.bits, self.num_copies, self.num_extra_constants = bits, num_copies, num_extra_constants
        self.params = (bits, num_copies, num_extra_constants)

    @classmethod
    def new_from_config(cls, num_wires, num_routed_wires, num_constants, bits):
        vec = 1 << bits
        copies = min(num_routed_wires // (2 + vec), num_wires // (2 + vec + bits))
        extra = num_routed_wires - (2 + vec) * copies
        return cls(bits, copies, min(extra, num_constants))

    def vec_size(self):
        return 1 << self.bits

    def num_routed(self):
        return (2 + self.vec_size()) * self.num_copies + self.num_extra_constants

    def num_constraints(self):
        return (self.bits + 2) * self.num_copies + self.num_extra_constants

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        vs = self.vec_size()
        for copy in range(self.num_copies):
            base = (2 + vs) * copy
            access_index, claimed = wires[base], wires[base + 1]
            items = [wires[base + 2 + i] for i in range(vs)]
            bits = [wires[self.num_routed() + copy * self.bits + i] for i in range(self.bits)]
            for b in bits:
                out.append(b * (b - 1) % P)
            acc = 0
            for b in reversed(bits):
                acc = (acc + acc + b) % P
            out.append((acc - access_index) % P)
            for b in bits:
                items = [(items[2 * k] + b * (items[2 * k + 1] - items[2 * k])) % P for k in range(len(items) // 2)]
            out.append((items[0] - claimed) % P)
        start = (2 + vs) * self.num_copies
        for i in range(self.num_extra_constants):
            out.append((consts[i] - wires[start + i]) % P)
        return out


def _limb_product(limb, max_limb=4):
    p = 1
    for x in range(max_limb):
        p = p * (limb - x) % P
    return p


class U32ArithmeticGate(Gate):
    """u32/src/gates/arithmetic_u32.rs:326-386"""
    type_id = 7

    def __init__(self, num_ops):
        self.num_ops, self.params = num_ops, (num_ops,)

    @staticmethod
    def num_ops_for(num_wires, num_routed):
        return min(num_wires // (6 + 32), num_routed // 6)

    def num_constraints(self):
        return self.num_ops * (4 + 32)

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        for i in range(self.num_ops):
            m0, m1, add, lo, hi, inverse = wires[6 * i:6 * i + 6]
            computed = (m0 * m1 + add) % P
            diff = (0xFFFFFFFF - hi) % P
            hi_not_max = (inverse * diff - 1) % P
            out.append(hi_not_max * lo % P)
            out.append((hi * (1 << 32) + lo - computed) % P)
            cl = ch = 0
            for j in reversed(range(32)):
                limb = wires[6 * self.num_ops + 32 * i + j]
                out.append(_limb_product(limb))
                if j < 16:
                    cl = (cl * 4 + limb) % P
                else:
                    ch = (ch * 4 + limb) % P
            out.append((cl - lo) % P)
            out.append((ch - hi) % P)
        return out

    def honest_op(self, m0, m1, add):
        """wires of one op as U32ArithmeticGenerator sets them (arithmetic_u32.rs:420-470)"""
        v = m0 * m1 + add
        lo, hi = v & 0xFFFFFFFF, v >> 32
        inverse = inv((0xFFFFFFFF - hi) % P) if hi != 0xFFFFFFFF else 0
        limbs = [(v >> (2 * j)) & 3 for j in range(32)]
        return [m0, m1, add, lo, hi, inverse], limbs


class U32AddManyGate(Gate):
    """u32/src/gates/add_many_u32.rs:143-184"""
    type_id = 8

    def __init__(self, num_addends, num_ops):
        self.num_addends, self.num_ops, self.params = num_addends, num_ops, (num_addends, num_ops)

    @staticmethod
    def num_ops_for(num_addends, num_wires, num_routed):
        return min(num_wires // (num_addends + 3 + 18), num_routed // (num_addends + 3))

    def num_constraints(self):
        return self.num_ops * (3 + 18)

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        na = self.num_addends
        for i in range(self.num_ops):
            b = (na + 3) * i
            addends = wires[b:b + na]
            carry, res, ocarry = wires[b + na], wires[b + na + 1], wires[b + na + 2]
            computed = (sum(addends) + carry) % P
            out.append((ocarry * (1 << 32) + res - computed) % P)
            cr = cc = 0
            for j in reversed(range(18)):
                limb = wires[(na + 3) * self.num_ops + 18 * i + j]
                out.append(_limb_product(limb))
                if j < 16:
                    cr = (4 * cr + limb) % P
                else:
                    cc = (4 * cc + limb) % P
            out.append((cr - res) % P)
            out.append((cc - ocarry) % P)
        return out


class U32RangeCheckGate(Gate):
    """u32/src/gates/range_check_u32.rs:89-111"""
    type_id = 9

    def __init__(self, num_input_limbs):
        self.num_input_limbs, self.params = num_input_limbs, (num_input_limbs,)

    def num_constraints(self):
        return self.num_input_limbs * 17

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        n = self.num_input_limbs
        for i in range(n):
            aux = wires[n + 16 * i:n + 16 * i + 16]
            out.append((reduce_with_powers(aux, 4) - wires[i]) % P)
            for a in aux:
                out.append(_limb_product(a))
        return out


class U32SubtractionGate(Gate):
    """u32/src/gates/subtraction_u32.rs:233-271"""
    type_id = 10

    def __init__(self, num_ops):
        self.num_ops, self.params = num_ops, (num_ops,)

    @staticmethod
    def num_ops_for(num_wires, num_routed):
        return min(num_wires // (5 + 16), num_routed // 5)

    def num_constraints(self):
        return self.num_ops * (3 + 16)

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        for i in range(self.num_ops):
            x, y, borrow, res, oborrow = wires[5 * i:5 * i + 5]
            initial = (x - y - borrow) % P
            out.append((res - (initial + oborrow * (1 << 32))) % P)
            comb = 0
            for j in reversed(range(16)):
                limb = wires[5 * self.num_ops + 16 * i + j]
                out.append(_limb_product(limb))
                comb = (comb * 4 + limb) % P
            out.append((comb - res) % P)
            out.append(oborrow * (1 - oborrow) % P)
        return out


class ComparisonGate(Gate):
    """u32/src/gates/comparison.rs:325-402"""
    type_id = 11

    def __init__(self, num_bits, num_chunks):
        self.num_bits, self.num_chunks, self.params = num_bits, num_chunks, (num_bits, num_chunks)

    def chunk_bits(self):
        return -(-self.num_bits // self.num_chunks)

    def num_constraints(self):
        return 6 + 5 * self.num_chunks + self.chunk_bits()

    def eval_unfiltered(self, consts, wires, pih):
        nc, cb = self.num_chunks, self.chunk_bits()
        first, second = wires[0], wires[1]
        fc = wires[4:4 + nc]
        sc = wires[4 + nc:4 + 2 * nc]
        out = [(reduce_with_powers(fc, 1 << cb) - first) % P, (reduce_with_powers(sc, 1 << cb) - second) % P]
        msd = 0
        for i in range(nc):
            fp = sp = 1
            for x in range(1 << cb):
                fp = fp * (fc[i] - x) % P
                sp = sp * (sc[i] - x) % P
            out += [fp, sp]
            diff = (sc[i] - fc[i]) % P
            dummy, eq, inter = wires[4 + 2 * nc + i], wires[4 + 3 * nc + i], wires[4 + 4 * nc + i]
            out.append((diff * dummy - (1 - eq)) % P)
            out.append(eq * diff % P)
            out.append((inter - eq * msd) % P)
            msd = (inter + (1 - eq) * diff) % P
        out.append((wires[3] - msd) % P)
        bits = wires[4 + 5 * nc:4 + 5 * nc + cb + 1]
        for b in bits:
            out.append(b * (1 - b) % P)
        out.append((wires[3] + (1 << cb) - reduce_with_powers(bits, 2)) % P)
        out.append((wires[2] - bits[cb]) % P)
        return out


GATE_TYPES = {c.type_id: c for c in (NoopGate, ConstantGate, PublicInputGate, ArithmeticGate, BaseSumGate, PoseidonGate,
                                     RandomAccessGate, U32ArithmeticGate, U32AddManyGate, U32RangeCheckGate, U32SubtractionGate,
                                     ComparisonGate)}


# ---------------------------------------------------------------------------------------------------------------
# circuit description + vanishing polynomial
# ---------------------------------------------------------------------------------------------------------------
class Circuit:
    """The part of CommonCircuitData the quotient evaluation reads (plonk/circuit_data.rs:270-349)."""

    def __init__(self, gates, selector_indices, groups, num_wires, num_routed_wires, num_constants, k_is, degree_bits,
                 rate_bits=3, num_challenges=2, quotient_degree_factor=8):
        self.gates, self.selector_indices, self.groups = gates, selector_indices, groups
        self.num_wires, self.num_routed_wires, self.num_constants = num_wires, num_routed_wires, num_constants
        self.k_is, self.degree_bits, self.rate_bits = [k % P for k in k_is], degree_bits, rate_bits
        self.num_challenges, self.quotient_degree_factor = num_challenges, quotient_degree_factor
        self.num_selectors = len(groups)
        self.num_gate_constraints = max([g.num_constraints() for g in gates] + [0])
        self.num_partial_products = -(-num_routed_wires // quotient_degree_factor) - 1  # partial_products.rs:40-47
        self.quotient_degree_bits = (quotient_degree_factor - 1).bit_length()          # log2_ceil
        assert self.quotient_degree_bits <= rate_bits                                    # prover.rs:809-813


def compute_filter(row, group, s, many_selectors):
    """gates/gate.rs:261-268"""
    f = 1
    for i in range(group[0], group[1]):
        if i != row:
            f = f * (i - s) % P
    if many_selectors:
        f = f * (UNUSED_SELECTOR - s) % P
    return f


def evaluate_gate_constraints(circ, consts, wires, pih):
    """vanishing_poly.rs:267-306 for one point: constraint j = sum over gates of filter * c_{g,j}"""
    acc = [0] * circ.num_gate_constraints
    for i, gate in enumerate(circ.gates):
        si = circ.selector_indices[i]
        filt = compute_filter(i, circ.groups[si], consts[si], circ.num_selectors > 1)
        cs = gate.eval_unfiltered(consts[circ.num_selectors:], wires, pih)
        assert len(cs) == gate.num_constraints()
        for j, c in enumerate(cs):
            acc[j] = (acc[j] + filt * c) % P
    return acc


def check_partial_products(nums, dens, partials, z_x, z_gx, max_degree):
    """util/partial_products.rs:52-78"""
    accs = [z_x] + list(partials) + [z_gx]
    out = []
    for c in range(-(-len(nums) // max_degree)):
        np_, dp = 1, 1
        for v in nums[c * max_degree:(c + 1) * max_degree]:
            np_ = np_ * v % P
        for v in dens[c * max_degree:(c + 1) * max_degree]:
            dp = dp * v % P
        out.append((accs[c] * np_ - accs[c + 1] * dp) % P)
    return out


def wires_permutation_partial_products_and_zs(wires_values, sigma_values, k_is, beta, gamma, max_degree, degree_bits):
    """plonk/prover.rs:729-786 for one challenge.  wires_values [num_wires][n], sigma_values [num_routed][n] (values on H,
    row i at x = w^i).  Returns [num_partial_products + 1][n]: the partial-product polynomials, then Z LAST (the caller
    pops it to the front, prover.rs:112-117)."""
    n = 1 << degree_bits
    nr = len(k_is)
    w = root_of_unity(degree_bits)
    x, z_x, rows = 1, 1, []
    for i in range(n):
        nums = [(int(wires_values[j][i]) + beta * (k_is[j] * x % P) + gamma) % P for j in range(nr)]
        dens = [(int(wires_values[j][i]) + beta * int(sigma_values[j][i]) + gamma) % P for j in range(nr)]
        quot = [a * inv(b) % P for a, b in zip(nums, dens)]       # batch_multiplicative_inverse + mul, :757-761
        chunk = []                                               # quotient_chunk_products, partial_products.rs:13-24
        for c in range(0, nr, max_degree):
            v = 1
            for q in quot[c:c + max_degree]:
                v = v * q % P
            chunk.append(v)
        acc, res = z_x, []                                       # partial_products_and_z_gx, :28-37
        for v in chunk:
            acc = acc * v % P
            res.append(acc)
        z_x, res[-1] = res[-1], z_x                              # swap(&mut z_x, &mut res[num_prods]), :778
        rows.append(res)
        x = x * w % P
    return [[rows[i][k] for i in range(n)] for k in range(len(rows[0]))]   # transpose


def all_zs_partial_products(wires_values, sigma_values, k_is, betas, gammas, max_degree, degree_bits):
    """prover.rs:106-117: the matrix committed as zs_partial_products: [Z_c for c] + [pp_{c,k} for c for k]."""
    per = [wires_permutation_partial_products_and_zs(wires_values, sigma_values, k_is, b, g, max_degree, degree_bits)
           for b, g in zip(betas, gammas)]
    zs = [pp.pop() for pp in per]
    return zs + [col for pp in per for col in pp]


def root_of_unity(k):
    b = 1753635133440165772
    for _ in range(32 - k):
        b = b * b % P
    return b


def quotient_point_rows(circ, i):
    """Leaf indices (row of x_i, row of g * x_i's `next` point) that point i of the quotient domain reads (prover.rs:923-943)."""
    n_log, qdb, rb = circ.degree_bits, circ.quotient_degree_bits, circ.rate_bits
    lde_size, lde_bits = 1 << (n_log + qdb), n_log + rb
    step, next_step = 1 << (rb - qdb), 1 << qdb
    rev = lambda v: int(format(v, "0%db" % lde_bits)[::-1], 2) if lde_bits else 0
    return rev(i * step), rev(((i + next_step) % lde_size) * step)


def compute_quotient_values(circ, wires_rows, zs_pp_rows, consts_sigmas_rows, pih, betas, gammas, alphas, points=None):
    """The loop of prover.rs:884-999: rows are the three batches' LDE leaves (leaf order, as the commit produces
    them); returns quotient_values[i][challenge] for i over the 2^(degree_bits + quotient_degree_bits) points -- or only
    for the listed `points` (the row containers then only need the rows quotient_point_rows() names: full-size parity on a
    sample, tests/test_gpu_configs.py)."""
    n_log, qdb, rb = circ.degree_bits, circ.quotient_degree_bits, circ.rate_bits
    lde_size = 1 << (n_log + qdb)
    step, next_step = 1 << (rb - qdb), 1 << qdb
    lde_bits = n_log + rb

    def rev(i):
        return int(format(i, "0%db" % lde_bits)[::-1], 2) if lde_bits else 0

    w = root_of_unity(n_log + qdb)
    g = 7
    # ZeroPolyOnCoset::new (zero_poly_coset.rs:20-33)
    g_pow_n = pow(g, 1 << n_log, P)
    v = root_of_unity(qdb)
    zh = [(g_pow_n * pow(v, i, P) - 1) % P for i in range(1 << qdb)]
    zh_inv = [inv(z) for z in zh]
    nc, nr, md, npp = circ.num_challenges, circ.num_routed_wires, circ.quotient_degree_factor, circ.num_partial_products
    out = []
    for i in (range(lde_size) if points is None else points):
        x = g * pow(w, i, P) % P
        row = rev(i * step)
        row_next = rev(((i + next_step) % lde_size) * step)
        cs = [int(t) for t in consts_sigmas_rows[row]]
        consts, sigmas = cs[:circ.num_constants], cs[circ.num_constants:circ.num_constants + nr]
        wires = [int(t) for t in wires_rows[row]][:circ.num_wires]
        zpp = [int(t) for t in zs_pp_rows[row]]
        zs, pps = zpp[:nc], zpp[nc:nc * (1 + npp)]
        next_zs = [int(t) for t in zs_pp_rows[row_next]][:nc]
        gate_terms = evaluate_gate_constraints(circ, consts, wires, pih)
        l0 = zh[i % (1 << qdb)] * inv((1 << n_log) * (x - 1)) % P     # eval_l_0, zero_poly_coset.rs:57-60
        z1_terms, pp_terms = [], []
        for c in range(nc):
            z1_terms.append(l0 * (zs[c] - 1) % P)
            nums = [(wires[j] + betas[c] * (circ.k_is[j] * x % P) + gammas[c]) % P for j in range(nr)]
            dens = [(wires[j] + betas[c] * sigmas[j] + gammas[c]) % P for j in range(nr)]
            pp_terms += check_partial_products(nums, dens, pps[c * npp:(c + 1) * npp], zs[c], next_zs[c], md)
        terms = z1_terms + pp_terms + gate_terms
        res = []
        for a in alphas:                                              # reduce_with_powers_multi
            acc = 0
            for t in reversed(terms):
                acc = (t + acc * a) % P
            res.append(acc * zh_inv[i % (1 << qdb)] % P)              # prover.rs:985-991
        out.append(res)
    return out


def compute_quotient_polys(circ, *args):
    """prover.rs:1009-1021: transpose, then coset_ifft(g) per challenge.  Returns (values, coefficients)."""
    import numpy as np
    from . import coset_ifft
    vals = compute_quotient_values(circ, *args)
    cols = [np.array([v[c] for v in vals], dtype=np.uint64) for c in range(circ.num_challenges)]
    return cols, [coset_ifft(c, 7) for c in cols]
