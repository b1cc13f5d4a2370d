// quotient.cuh -- quotient-polynomial evaluation over the LDE domain for any gate set.
//
// Reference (CPU):  compute_quotient_polys                 plonky2/src/plonk/prover.rs:790-1034
//                   eval_vanishing_poly_base_batch         plonky2/src/plonk/vanishing_poly.rs:100-226
//                   evaluate_gate_constraints_base_batch   vanishing_poly.rs:267-306, gates/gate.rs:113-150, 261-268
//                   check_partial_products                 util/partial_products.rs:52-78
//                   reduce_with_powers_multi               plonk/plonk_common.rs:97-114
//                   ZeroPolyOnCoset                        field/src/zero_poly_coset.rs:7-60
// Reference (GPU):  compute_quotient_values_kernel, cuda/plonky2_gpu_impl.cuh:485-878 -- hard-coded to one ed25519
//                   circuit (25 literal gate instances, 231 constraints, literal public-input hash); this kernel takes
//                   the circuit as data (p2b_circuit) instead.
//
// One thread per LDE point i (x = g * w^i).  It reads its three leaf rows (leaf index = reverse_bits(i * step), the
// layout the commit kernels produce), evaluates every gate's constraints, and folds each term straight into the
// per-challenge sums  sum_t alpha_c^t * term_t  with a table of alpha powers -- the same value as the reference's
// Horner pass from the last term (exact field arithmetic), without the 231-word per-thread constraint array the
// reference keeps in local memory:  gate constraint j of gate g contributes filter_g * alpha^(T0 + j) * c_{g,j}, so a
// gate accumulates sum_j alpha^(T0+j) c_{g,j} and is multiplied by its filter once.
#pragma once
#include "poseidon.cuh"

namespace quotient {

using gl::u32;
using gl::u64;

enum GateType : u32 {
  G_NOOP = 0, G_CONSTANT = 1, G_PUBLIC_INPUT = 2, G_ARITHMETIC = 3, G_BASE_SUM = 4, G_POSEIDON = 5, G_RANDOM_ACCESS = 6,
  G_U32_ARITHMETIC = 7, G_U32_ADD_MANY = 8, G_U32_RANGE_CHECK = 9, G_U32_SUBTRACTION = 10, G_COMPARISON = 11,
  // the gates a recursive-verifier circuit adds (extension elements = 2 consecutive wires)
  G_ARITHMETIC_EXT = 12, G_MUL_EXT = 13, G_REDUCING = 14, G_REDUCING_EXT = 15, G_EXPONENTIATION = 16, G_POSEIDON_MDS = 17,
  G_HIGH_DEGREE_INTERPOLATION = 18, G_LOW_DEGREE_INTERPOLATION = 19, G_NUM_TYPES = 20
};

struct GateDesc {  // == p2b_gate
  u32 type, selector_index, group_start, group_end, p0, p1, p2, reserved;
};

static constexpr int MAX_CHALLENGES = 4;
static constexpr int MAX_ZH = 32;

struct Params {
  u32 degree_bits, rate_bits, qdb, num_challenges;
  u32 num_wires, num_routed, num_constants, num_selectors, num_partial_products, max_degree, num_gates, num_gate_constraints;
  const u64* wires; u64 wires_stride;   // leaf rows of the three commitments
  const u64* zs_pp; u64 zs_stride;
  const u64* cs; u64 cs_stride;
  const GateDesc* gates;                // device
  const u64* k_is;                      // device [num_routed]
  const u64* alpha_pows;                // device [num_challenges][num_terms]
  u32 num_terms;
  u64 pih[4];
  u64 betas[MAX_CHALLENGES], gammas[MAX_CHALLENGES];
  u64 zh[MAX_ZH], zh_inv[MAX_ZH];       // Z_H on the coset and inverses, indexed i mod 2^qdb
  u64 w;                                // generator of the evaluation subgroup (order 2^(degree_bits + qdb))
  u64 n_field;                          // n as a field element
  u64 small_roots[9];                   // primitive_root_of_unity(k), k <= 8 (interpolation gates' cosets)
  u64* out_values;                      // [num_challenges][pt_count]
  u64* out_rows;                        // optional [pt_count][num_challenges] (the reference's d_outs layout)
  // the points this launch evaluates: i = pt_first + pt_stride * j, j < pt_count (the whole domain: 0, 1, lde_size).  A device
  // of a multi-device prover owns the points whose rows lie in its leaf range: i = rev(d) mod G (csrc/mgpu.cuh).
  u64 pt_first, pt_stride, pt_count;
  u64 last_row;                         // index of the last leaf row the three matrices hold (copied without over-read)
};

// host: number of constraints of a gate (each gate's num_constraints())
__host__ __device__ inline u32 gate_num_constraints(const GateDesc& g) {
  switch (g.type) {
    case G_NOOP: return 0;
    case G_CONSTANT: return g.p0;
    case G_PUBLIC_INPUT: return 4;
    case G_ARITHMETIC: return g.p0;
    case G_BASE_SUM: return 1 + g.p0;
    case G_POSEIDON: return 123;
    case G_RANDOM_ACCESS: return (g.p0 + 2) * g.p1 + g.p2;
    case G_U32_ARITHMETIC: return g.p0 * 36;
    case G_U32_ADD_MANY: return g.p1 * 21;
    case G_U32_RANGE_CHECK: return g.p0 * 17;
    case G_U32_SUBTRACTION: return g.p0 * 19;
    case G_COMPARISON: return 6 + 5 * g.p1 + (g.p0 + g.p1 - 1) / g.p1;
    case G_ARITHMETIC_EXT: return 2 * g.p0;
    case G_MUL_EXT: return 2 * g.p0;
    case G_REDUCING: return 2 * g.p0;
    case G_REDUCING_EXT: return 2 * g.p0;
    case G_EXPONENTIATION: return g.p0 + 1;
    case G_POSEIDON_MDS: return 24;
    case G_HIGH_DEGREE_INTERPOLATION: return 2 * (1u << g.p0) + 2;
    case G_LOW_DEGREE_INTERPOLATION: return ((1u << g.p0) - 2) + 2 * (1u << g.p0) + 2 * ((1u << g.p0) - 2) + 2;
    default: return 0;
  }
}

// Accumulates sum_j alpha_c^(base + j) * c_j for every challenge; emit() is called in constraint order.  M = exact or
// optimistic field reduction (gl64.cuh): the kernel evaluates a point optimistically and redoes it exactly if `rare`.
template <class M, int NC>
struct Emitter {
  M& mode;
  __device__ __forceinline__ explicit Emitter(M& m) : mode(m) {}
  __device__ __forceinline__ u64 mul(u64 a, u64 b) { return gl::mul(a, b, mode); }
  __device__ __forceinline__ u64 mul_add(u64 a, u64 b, u64 c) { return gl::mul_add(a, b, c, mode); }
  // prod_{x < m} (limb - x)
  __device__ __forceinline__ u64 range_product(u64 limb, u32 m) {
    u64 p = limb;
    for (u32 x = 1; x < m; x++) p = mul(p, gl::sub_canonical(limb, x));
    return p;
  }
  // sum_j v_j * alpha_c^(base + j) accumulated UNREDUCED in 160 bits per challenge (v_j < 2^64, powers canonical: a gate
  // has far fewer than 2^32 constraints) and reduced once per gate: a constraint costs a 64x64 multiply and a 3-word
  // add per challenge instead of a full modular multiply-add
  u64 lo[NC], hi[NC];
  u32 top[NC];
  const u64* ap;  // alpha_pows + base
  u32 stride, j;
  __device__ __forceinline__ void init(const u64* alpha_pows, u32 num_terms, u32 base) {
    ap = alpha_pows + base;
    stride = num_terms;
    j = 0;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      lo[c] = 0;
      hi[c] = 0;
      top[c] = 0;
    }
  }
  __device__ __forceinline__ void emit(u64 v) {
#pragma unroll
    for (int c = 0; c < NC; c++) poseidon::mac160(lo[c], hi[c], top[c], v, __ldg(ap + (u64)c * stride + j));
    j++;
  }
  __device__ __forceinline__ u64 result(int c) { return poseidon::reduce160(lo[c], hi[c], top[c], mode); }
};

__device__ __forceinline__ u64 fsub(u64 a, u64 b) { return gl::sub(a, b); }
__device__ __forceinline__ u64 fadd(u64 a, u64 b) { return gl::add(a, b); }

// ---- gates (wires: W[k], constants after the selector prefix: K[k]) ----------------------------------------------
#define W(k) (w[(k)])
#define K(k) (kc[(k)])

template <class E>
__device__ __forceinline__ void eval_constant(const GateDesc& g, const u64* w, const u64* kc, E& e) {  // constant.rs:150-158
  for (u32 i = 0; i < g.p0; i++) e.emit(fsub(K(i), W(i)));
}
template <class E>
__device__ __forceinline__ void eval_public_input(const u64* w, const u64* pih, E& e) {  // public_input.rs:129-139
  for (u32 i = 0; i < 4; i++) e.emit(fsub(W(i), pih[i]));
}
template <class E>
__device__ __forceinline__ void eval_arithmetic(const GateDesc& g, const u64* w, const u64* kc, E& e) {  // arithmetic_base.rs:199-220
  u64 c0 = K(0), c1 = K(1);
  for (u32 i = 0; i < g.p0; i++) {
    u64 computed = fadd(e.mul(e.mul(W(4 * i), W(4 * i + 1)), c0), e.mul(W(4 * i + 2), c1));
    e.emit(fsub(W(4 * i + 3), computed));
  }
}
template <class E>
__device__ __forceinline__ void eval_base_sum(const GateDesc& g, const u64* w, E& e) {  // base_sum.rs:213-230
  const u32 nl = g.p0, B = g.p1;
  u64 sum = 0;
  for (u32 i = nl; i-- > 0;) sum = fadd(e.mul(sum, B), W(1 + i));
  e.emit(fsub(sum, W(0)));
  for (u32 i = 0; i < nl; i++) e.emit(e.range_product(W(1 + i), B));
}
template <class E>
__device__ __forceinline__ void eval_random_access(const GateDesc& g, const u64* w, const u64* kc, E& e) {  // random_access.rs:409-450
  const u32 bits = g.p0, copies = g.p1, extra = g.p2, vs = 1u << g.p0;
  const u32 routed = (2 + vs) * copies + extra;
  for (u32 copy = 0; copy < copies; copy++) {
    const u32 base = (2 + vs) * copy, bbase = routed + copy * bits;
    for (u32 i = 0; i < bits; i++) {
      u64 b = W(bbase + i);
      e.emit(e.mul(b, fsub(b, 1)));
    }
    u64 acc = 0;
    for (u32 i = bits; i-- > 0;) acc = fadd(fadd(acc, acc), W(bbase + i));
    e.emit(fsub(acc, W(base)));
    // fold the list: the item selected by the bits (bit 0 first, as the reference folds pairs)
    // items at level 0 are the wires; evaluate the selection tree recursively without a buffer (bits <= 6)
    u64 items[64];
    for (u32 i = 0; i < vs; i++) items[i] = W(base + 2 + i);
    u32 len = vs;
    for (u32 i = 0; i < bits; i++) {
      u64 b = W(bbase + i);
      len >>= 1;
      for (u32 k = 0; k < len; k++) items[k] = fadd(items[2 * k], e.mul(b, fsub(items[2 * k + 1], items[2 * k])));
    }
    e.emit(fsub(items[0], W(base + 1)));
  }
  for (u32 i = 0; i < extra; i++) e.emit(fsub(K(i), W((2 + vs) * copies + i)));
}
template <class E>
__device__ __forceinline__ void eval_u32_arithmetic(const GateDesc& g, const u64* w, E& e) {  // arithmetic_u32.rs:326-386
  const u32 ops = g.p0;
  for (u32 i = 0; i < ops; i++) {
    u64 m0 = W(6 * i), m1 = W(6 * i + 1), add = W(6 * i + 2), lo = W(6 * i + 3), hi = W(6 * i + 4), inverse = W(6 * i + 5);
    u64 computed = e.mul_add(m0, m1, add);
    u64 diff = fsub(0xFFFFFFFFull, hi);
    u64 hi_not_max = fsub(e.mul(inverse, diff), 1);
    e.emit(e.mul(hi_not_max, lo));
    e.emit(fsub(fadd(e.mul(hi, 1ull << 32), lo), computed));
    u64 cl = 0, ch = 0;
    for (u32 j = 32; j-- > 0;) {
      u64 limb = W(6 * ops + 32 * i + j);
      e.emit(e.range_product(limb, 4));
      if (j < 16) cl = fadd(e.mul(cl, 4), limb);
      else ch = fadd(e.mul(ch, 4), limb);
    }
    e.emit(fsub(cl, lo));
    e.emit(fsub(ch, hi));
  }
}
template <class E>
__device__ __forceinline__ void eval_u32_add_many(const GateDesc& g, const u64* w, E& e) {  // add_many_u32.rs:143-184
  const u32 na = g.p0, ops = g.p1;
  for (u32 i = 0; i < ops; i++) {
    const u32 b = (na + 3) * i;
    u64 computed = 0;
    for (u32 j = 0; j < na; j++) computed = fadd(computed, W(b + j));
    computed = fadd(computed, W(b + na));
    u64 res = W(b + na + 1), oc = W(b + na + 2);
    e.emit(fsub(fadd(e.mul(oc, 1ull << 32), res), computed));
    u64 cr = 0, cc = 0;
    for (u32 j = 18; j-- > 0;) {
      u64 limb = W((na + 3) * ops + 18 * i + j);
      e.emit(e.range_product(limb, 4));
      if (j < 16) cr = fadd(e.mul(cr, 4), limb);
      else cc = fadd(e.mul(cc, 4), limb);
    }
    e.emit(fsub(cr, res));
    e.emit(fsub(cc, oc));
  }
}
template <class E>
__device__ __forceinline__ void eval_u32_range_check(const GateDesc& g, const u64* w, E& e) {  // range_check_u32.rs:89-111
  const u32 n = g.p0;
  for (u32 i = 0; i < n; i++) {
    u64 sum = 0;
    for (u32 j = 16; j-- > 0;) sum = fadd(e.mul(sum, 4), W(n + 16 * i + j));
    e.emit(fsub(sum, W(i)));
    for (u32 j = 0; j < 16; j++) e.emit(e.range_product(W(n + 16 * i + j), 4));
  }
}
template <class E>
__device__ __forceinline__ void eval_u32_subtraction(const GateDesc& g, const u64* w, E& e) {  // subtraction_u32.rs:233-271
  const u32 ops = g.p0;
  for (u32 i = 0; i < ops; i++) {
    u64 x = W(5 * i), y = W(5 * i + 1), br = W(5 * i + 2), res = W(5 * i + 3), ob = W(5 * i + 4);
    u64 initial = fsub(fsub(x, y), br);
    e.emit(fsub(res, fadd(initial, e.mul(ob, 1ull << 32))));
    u64 comb = 0;
    for (u32 j = 16; j-- > 0;) {
      u64 limb = W(5 * ops + 16 * i + j);
      e.emit(e.range_product(limb, 4));
      comb = fadd(e.mul(comb, 4), limb);
    }
    e.emit(fsub(comb, res));
    e.emit(e.mul(ob, fsub(1, ob)));
  }
}
template <class E>
__device__ __forceinline__ void eval_comparison(const GateDesc& g, const u64* w, E& e) {  // comparison.rs:325-402
  const u32 nc = g.p1, cb = (g.p0 + g.p1 - 1) / g.p1;
  u64 fcomb = 0, scomb = 0;
  for (u32 i = nc; i-- > 0;) {
    fcomb = fadd(e.mul(fcomb, 1ull << cb), W(4 + i));
    scomb = fadd(e.mul(scomb, 1ull << cb), W(4 + nc + i));
  }
  e.emit(fsub(fcomb, W(0)));
  e.emit(fsub(scomb, W(1)));
  u64 msd = 0;
  for (u32 i = 0; i < nc; i++) {
    u64 f = W(4 + i), s = W(4 + nc + i);
    e.emit(e.range_product(f, 1u << cb));
    e.emit(e.range_product(s, 1u << cb));
    u64 diff = fsub(s, f);
    u64 dummy = W(4 + 2 * nc + i), eq = W(4 + 3 * nc + i), inter = W(4 + 4 * nc + i);
    e.emit(fsub(e.mul(diff, dummy), fsub(1, eq)));
    e.emit(e.mul(eq, diff));
    e.emit(fsub(inter, e.mul(eq, msd)));
    msd = fadd(inter, e.mul(fsub(1, eq), diff));
  }
  e.emit(fsub(W(3), msd));
  u64 bits_comb = 0;
  for (u32 i = 0; i <= cb; i++) {
    u64 b = W(4 + 5 * nc + i);
    e.emit(e.mul(b, fsub(1, b)));
  }
  for (u32 i = cb + 1; i-- > 0;) bits_comb = fadd(fadd(bits_comb, bits_comb), W(4 + 5 * nc + i));
  e.emit(fsub(fadd(W(3), 1ull << cb), bits_comb));
  e.emit(fsub(W(2), W(4 + 5 * nc + cb)));
}
// ---- recursion gate set: quadratic-extension arithmetic on wire pairs (field/src/extension/quadratic.rs:172-185) ----
struct X2 {
  u64 a, b;
};
#define WX(k) X2{W(k), W((k) + 1)}
__device__ __forceinline__ X2 xadd(X2 x, X2 y) { return X2{fadd(x.a, y.a), fadd(x.b, y.b)}; }
__device__ __forceinline__ X2 xsub(X2 x, X2 y) { return X2{fsub(x.a, y.a), fsub(x.b, y.b)}; }
template <class E>
__device__ __forceinline__ X2 xmul(X2 x, X2 y, E& e) {
  u64 t = e.mul(x.b, y.b);
  return X2{e.mul_add(x.a, y.a, e.mul(7, t)), e.mul_add(x.a, y.b, e.mul(x.b, y.a))};
}
template <class E>
__device__ __forceinline__ X2 xscale(X2 x, u64 s, E& e) { return X2{e.mul(x.a, s), e.mul(x.b, s)}; }
template <class E>
__device__ __forceinline__ void emit2(X2 v, E& e) {  // yield_constr.many(x.to_basefield_array())
  e.emit(v.a);
  e.emit(v.b);
}
template <class E>
__device__ __forceinline__ void eval_arithmetic_ext(const GateDesc& g, const u64* w, const u64* kc, E& e) {  // arithmetic_extension.rs:129-147
  const u64 c0 = K(0), c1 = K(1);
  for (u32 i = 0; i < g.p0; i++) {
    X2 computed = xadd(xscale(xmul(WX(8 * i), WX(8 * i + 2), e), c0, e), xscale(WX(8 * i + 4), c1, e));
    emit2(xsub(WX(8 * i + 6), computed), e);
  }
}
template <class E>
__device__ __forceinline__ void eval_mul_ext(const GateDesc& g, const u64* w, const u64* kc, E& e) {  // multiplication_extension.rs:122-137
  const u64 c0 = K(0);
  for (u32 i = 0; i < g.p0; i++) emit2(xsub(WX(6 * i + 4), xscale(xmul(WX(6 * i), WX(6 * i + 2), e), c0, e)), e);
}
// reducing.rs:160-180 (EXT = false: base-field coefficients, one wire each) / reducing_extension.rs:160-179 (EXT = true)
template <bool EXT, class E>
__device__ __forceinline__ void eval_reducing(const GateDesc& g, const u64* w, E& e) {
  const u32 n = g.p0, cw = EXT ? 2 : 1, start_accs = 6 + cw * n;
  const X2 alpha = WX(2);
  X2 acc = WX(4);
  for (u32 i = 0; i < n; i++) {
    const X2 coeff = EXT ? WX(6 + 2 * i) : X2{W(6 + i), 0};
    const X2 a = i + 1 == n ? WX(0) : WX(start_accs + 2 * i);  // the last accumulator is the output
    emit2(xsub(xadd(xmul(acc, alpha, e), coeff), a), e);
    acc = a;
  }
}
template <class E>
__device__ __forceinline__ void eval_exponentiation(const GateDesc& g, const u64* w, E& e) {  // exponentiation.rs:266-299
  const u32 nb = g.p0;
  const u64 base = W(0);
  u64 prev_inter = 0;
  for (u32 i = 0; i < nb; i++) {
    const u64 prev = i == 0 ? 1 : e.mul(prev_inter, prev_inter);
    const u64 bit = W(1 + (nb - 1 - i));  // little-endian bits, accumulated from the top
    const u64 inter = W(2 + nb + i);
    e.emit(fsub(e.mul(prev, fadd(e.mul(bit, base), fsub(1, bit))), inter));
    prev_inter = inter;
  }
  e.emit(fsub(W(1 + nb), prev_inter));
}
template <class E>
__device__ __forceinline__ void eval_poseidon_mds(const u64* w, E& e) {  // poseidon_mds.rs:184-203
  for (u32 r = 0; r < 12; r++) {
    X2 acc{0, 0};
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const u32 k = 2 * ((i + r) % 12);
      acc.a = e.mul_add(W(k), (u64)poseidon::mds_circ(i), acc.a);
      acc.b = e.mul_add(W(k + 1), (u64)poseidon::mds_circ(i), acc.b);
    }
    if (r == 0) {  // MDS_MATRIX_DIAG = [8, 0, ...]
      acc.a = e.mul_add(W(0), (u64)poseidon::MDS_DIAG0, acc.a);
      acc.b = e.mul_add(W(1), (u64)poseidon::MDS_DIAG0, acc.b);
    }
    emit2(xsub(WX(24 + 2 * r), acc), e);
  }
}
// gates/interpolation.rs:21-76 layout: shift 0 | values 1.. | evaluation point | evaluation value | coefficients
template <class E>
__device__ __forceinline__ void eval_high_degree_interpolation(const GateDesc& g, const u64* w, const u64* roots, E& e) {  // high_degree_interpolation.rs:126-147
  const u32 np = 1u << g.p0, ep_w = 1 + 2 * np, ev_w = ep_w + 2, coeffs_w = ev_w + 2;
  const u64 gen = roots[g.p0];
  u64 point = W(0);
  for (u32 i = 0; i < np; i++) {
    X2 acc{0, 0};
    for (u32 k = np; k-- > 0;) acc = xadd(xscale(acc, point, e), WX(coeffs_w + 2 * k));  // eval_base
    emit2(xsub(WX(1 + 2 * i), acc), e);
    point = e.mul(point, gen);
  }
  const X2 ep = WX(ep_w);
  X2 acc{0, 0};
  for (u32 k = np; k-- > 0;) acc = xadd(xmul(acc, ep, e), WX(coeffs_w + 2 * k));
  emit2(xsub(WX(ev_w), acc), e);
}
template <class E>
__device__ __forceinline__ void eval_low_degree_interpolation(const GateDesc& g, const u64* w, const u64* roots, E& e) {  // low_degree_interpolation.rs:356-404
  const u32 np = 1u << g.p0, ep_w = 1 + 2 * np, ev_w = ep_w + 2, coeffs_w = ev_w + 2, end_coeffs = coeffs_w + 2 * np;
  auto shift_pow_w = [&](u32 i) { return i == 1 ? 0u : end_coeffs + i - 2; };                     // :50-57
  auto eval_pow_w = [&](u32 i) { return i == 1 ? ep_w : end_coeffs + np - 2 + 2 * (i - 2); };     // :59-67
  const u64 shift = W(0);
  for (u32 i = 1; i + 1 < np; i++) e.emit(fsub(e.mul(W(shift_pow_w(i)), shift), W(shift_pow_w(i + 1))));
  const u64 gen = roots[g.p0];
  u64 point = 1;
  for (u32 i = 0; i < np; i++) {
    X2 acc{0, 0};
    for (u32 k = np; k-- > 0;) {  // altered coefficient k = c_k * shift^k
      X2 c = WX(coeffs_w + 2 * k);
      if (k) c = xscale(c, W(shift_pow_w(k)), e);
      acc = xadd(xscale(acc, point, e), c);
    }
    emit2(xsub(WX(1 + 2 * i), acc), e);
    point = e.mul(point, gen);
  }
  const X2 ep = WX(ep_w);
  for (u32 i = 1; i + 1 < np; i++) emit2(xsub(xmul(WX(eval_pow_w(i)), ep, e), WX(eval_pow_w(i + 1))), e);
  X2 acc = WX(coeffs_w);  // eval_with_powers
  for (u32 k = 1; k < np; k++) acc = xadd(acc, xmul(WX(eval_pow_w(k)), WX(coeffs_w + 2 * k), e));
  emit2(xsub(WX(ev_w), acc), e);
}
// gates/poseidon.rs:485-564
template <class E>
__device__ __forceinline__ void eval_poseidon(const u64* w, E& e) {
  constexpr u32 SWAP = 24, DELTA = 25, FULL0 = 29, PARTIAL = 29 + 36, FULL1 = 29 + 36 + 22;
  u64 swap = W(SWAP);
  e.emit(e.mul(swap, fsub(swap, 1)));
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    u64 lhs = W(i), rhs = W(i + 4), d = W(DELTA + i);
    e.emit(fsub(e.mul(swap, fsub(rhs, lhs)), d));
    s[i] = fadd(lhs, d);
    s[i + 4] = fsub(rhs, d);
  }
#pragma unroll
  for (int i = 8; i < 12; i++) s[i] = W(i);
  using namespace poseidon;
  // The gate's constraints are stated in the reference through the fast partial-round form (gates/poseidon.rs:485-564); the
  // quantities they constrain -- the value entering every S-box, and the output state -- are the same in the naive round
  // structure (the fast form only changes the basis of lanes 1..11 between partial rounds, lane 0 is fixed by every one of
  // its matrices), so the evaluation walks the 30 naive rounds with the multiplier-free FP64 MDS layer of poseidon.cuh:
  // after mds_naive(.., constants of round r+1) the state IS the S-box input of round r+1.  Pinned against the reference's
  // own device evaluator on arbitrary rows (tests/test_ref_cuda_crosscheck.py).
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_canonical(s[i], C.rc[i]);
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    const bool full = r < 4 || r >= 26;
    if (full && r != 0) {
      const u32 base = r < 4 ? FULL0 + 12 * (r - 1) : FULL1 + 12 * (r - 26);
#pragma unroll 1
      for (int g4 = 0; g4 < 3; g4++) {
        // the state must equal the S-box input wires; processed through a 3 x 4 rotation to keep the register array
        // statically indexed
        u64 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3];
        u64 v0 = W(base + 4 * g4 + 0), v1 = W(base + 4 * g4 + 1), v2 = W(base + 4 * g4 + 2), v3 = W(base + 4 * g4 + 3);
        e.emit(fsub(t0, v0));
        e.emit(fsub(t1, v1));
        e.emit(fsub(t2, v2));
        e.emit(fsub(t3, v3));
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = s[i + 4];
        s[8] = v0; s[9] = v1; s[10] = v2; s[11] = v3;
      }
    }
    if (full) {
      sbox_layer_rolled(s, e.mode);
    } else {
      const u64 sbox_in = W(PARTIAL + (r - 4));
      e.emit(fsub(s[0], sbox_in));
      s[0] = sbox(sbox_in, e.mode);
    }
    mds_naive(s, C.rc_d + 24 * (r + 1), e.mode);
  }
#pragma unroll 1
  for (int g4 = 0; g4 < 3; g4++) {
    e.emit(fsub(s[0], W(12 + 4 * g4 + 0)));
    e.emit(fsub(s[1], W(12 + 4 * g4 + 1)));
    e.emit(fsub(s[2], W(12 + 4 * g4 + 2)));
    e.emit(fsub(s[3], W(12 + 4 * g4 + 3)));
    u64 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = s[i + 4];
    s[8] = t0; s[9] = t1; s[10] = t2; s[11] = t3;
  }
}
#undef W
#undef K

// field inverse by Fermat (the reference uses a binary GCD, field/src/inversion.rs; the value is the same)
__device__ __forceinline__ u64 finv(u64 a) { return gl::pow(a, gl::P - 2); }

// ---- the kernel ----------------------------------------------------------------------------------------------------
// A CTA of 12 warps owns a TILE of TW * 32 consecutive points of the quotient domain (TW = 1, 2, 3, 4 or 6 "point warps").
// Phase 1 stages the three leaf rows of every point of the tile (+ the next row's Z values) into shared memory with
// cp.async.bulk (one TMA bulk copy per row and matrix, completion on an mbarrier) -- a point's rows are contiguous 1-3 KB
// segments of the leaf-major matrices the commit kernels wrote, scattered by the bit reversal, which is exactly what a
// bulk copy wants and what per-thread loads are worst at (round 1: 12 warps per SM stalled on long-scoreboard loads 74% of
// the time, rows re-read from DRAM for every gate).  Phase 2: the 12 warps form 12 / TW groups of TW warps (one per point
// warp of the tile); a group pulls the next work item -- one gate instance or the permutation argument, ordered longest first
// by a host-side instruction-count model -- from a shared-memory queue and its warps evaluate that item for their 32 points
// each out of shared memory.  All lanes of a warp run the same evaluator (no divergence), the warps of a group share its
// instruction-cache lines, nothing waits on global memory, and the greedy longest-first order keeps the groups busy to the
// end whatever the model's errors (a static split left 2 stall cycles per issue at the final barrier).  Phase 3 adds the
// groups' partial sums, multiplies by 1 / Z_H and writes the values.  Every row is read from HBM exactly once.
struct PointRows {
  const u64 *w, *cs, *zp, *zn;
  u64 x;
};

// vanishing_z_1_terms and partial-product checks (vanishing_poly.rs:160-205) of one point.
// reduce_with_powers_multi reduces ONE term list [z1 terms of every challenge | pp checks of every challenge |
// gate constraints] with each alpha, so the terms produced for challenge tc enter every challenge c's sum.
template <class M, int NC>
__device__ __forceinline__ void eval_permutation_terms(const Params& p, const PointRows& r, const u64 i, M& mode, u64 (&acc)[NC]) {
  const u64 *w = r.w, *cs = r.cs, *zp = r.zp, *zn = r.zn;
  const u64 x = gl::mul(7, gl::pow(p.w, i));  // shifted_x = coset_shift * w^i (prover.rs:907)
  const u32 nr = p.num_routed, npp = p.num_partial_products, md = p.max_degree;
  const u32 rate_mask = (1u << p.qdb) - 1;
  auto fm = [&](u64 a, u64 b) { return gl::mul(a, b, mode); };
  auto fma = [&](u64 a, u64 b, u64 c) { return gl::mul_add(a, b, c, mode); };
#pragma unroll
  for (int c = 0; c < NC; c++) acc[c] = 0;
  // eval_l_0 (zero_poly_coset.rs:57-60)
  const u64 l0 = gl::mul(p.zh[i & rate_mask], finv(gl::mul(p.n_field, gl::sub(x, 1))));
  const u32 chunks = npp + 1;
#pragma unroll 1
  for (u32 tc = 0; tc < (u32)NC; tc++) {
    const u64 z_x = zp[tc], z_gx = zn[tc];
    const u64 t_z1 = fm(l0, gl::sub(z_x, 1));
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = fma(t_z1, __ldg(p.alpha_pows + (u64)c * p.num_terms + tc), acc[c]);
    const u64 beta = p.betas[tc], gamma = p.gammas[tc];
    const u64 bx = fm(beta, x);
    u64 prev = z_x;
    for (u32 ch = 0; ch < chunks; ch++) {
      u64 num = 1, den = 1;
      const u32 j1 = min(nr, (ch + 1) * md);
      for (u32 j = ch * md; j < j1; j++) {
        const u64 wv = w[j];
        num = fm(num, gl::add(gl::add(wv, fm(bx, __ldg(p.k_is + j))), gamma));               // :175-181
        den = fm(den, gl::add(gl::add(wv, fm(beta, cs[p.num_constants + j])), gamma));       // :182-186
      }
      const u64 next = ch + 1 < chunks ? zp[NC + tc * npp + ch] : z_gx;
      const u64 term = gl::sub(fm(prev, num), fm(next, den));  // partial_products.rs:70-75
#pragma unroll
      for (int c = 0; c < NC; c++)
        acc[c] = fma(term, __ldg(p.alpha_pows + (u64)c * p.num_terms + NC + tc * chunks + ch), acc[c]);
      prev = next;
    }
  }
}

// filter_g * sum_j alpha_c^(gate_base + j) c_{g,j} of one gate at one point (vanishing_poly.rs:267-306)
template <class M, int NC>
__device__ __forceinline__ void eval_gate_terms(const Params& p, const GateDesc& g, const u32 gi, const PointRows& r, M& mode,
                                                u64 (&out)[NC]) {
  const u64 *w = r.w, *cs = r.cs;
  const u64* kc = cs + p.num_selectors;  // vars.remove_prefix(num_selectors), gate.rs:138
  auto fm = [&](u64 a, u64 b) { return gl::mul(a, b, mode); };
  // compute_filter (gate.rs:261-268)
  const u64 s = cs[g.selector_index];
  u64 filter = 1;
  for (u32 k = g.group_start; k < g.group_end; k++)
    if (k != gi) filter = fm(filter, gl::sub((u64)k, s));
  if (p.num_selectors > 1) filter = fm(filter, gl::sub(0xFFFFFFFFull, s));
  Emitter<M, NC> e(mode);
  e.init(p.alpha_pows, p.num_terms, NC * (p.num_partial_products + 2));
  switch (g.type) {
    case G_NOOP: break;
    case G_CONSTANT: eval_constant(g, w, kc, e); break;
    case G_PUBLIC_INPUT: eval_public_input(w, p.pih, e); break;
    case G_ARITHMETIC: eval_arithmetic(g, w, kc, e); break;
    case G_BASE_SUM: eval_base_sum(g, w, e); break;
    case G_POSEIDON: eval_poseidon(w, e); break;
    case G_RANDOM_ACCESS: eval_random_access(g, w, kc, e); break;
    case G_U32_ARITHMETIC: eval_u32_arithmetic(g, w, e); break;
    case G_U32_ADD_MANY: eval_u32_add_many(g, w, e); break;
    case G_U32_RANGE_CHECK: eval_u32_range_check(g, w, e); break;
    case G_U32_SUBTRACTION: eval_u32_subtraction(g, w, e); break;
    case G_COMPARISON: eval_comparison(g, w, e); break;
    case G_ARITHMETIC_EXT: eval_arithmetic_ext(g, w, kc, e); break;
    case G_MUL_EXT: eval_mul_ext(g, w, kc, e); break;
    case G_REDUCING: eval_reducing<false>(g, w, e); break;
    case G_REDUCING_EXT: eval_reducing<true>(g, w, e); break;
    case G_EXPONENTIATION: eval_exponentiation(g, w, e); break;
    case G_POSEIDON_MDS: eval_poseidon_mds(w, e); break;
    case G_HIGH_DEGREE_INTERPOLATION: eval_high_degree_interpolation(g, w, p.small_roots, e); break;
    case G_LOW_DEGREE_INTERPOLATION: eval_low_degree_interpolation(g, w, p.small_roots, e); break;
    default: break;
  }
#pragma unroll
  for (int c = 0; c < NC; c++) out[c] = fm(filter, e.result(c));
}

static constexpr u32 QUOT_WARPS = 12;           // CTA = 384 threads
static constexpr u32 WORK_PERMUTATION = 0xFFFFFFFFu;  // work item that is not a gate: the permutation argument's terms

// Shared-memory geometry of a tile (u64 words).  A row slot holds [align pad <= 1][lead <= 1][row][tail pad <= 1]: the bulk
// copy's destination must be 16-byte aligned and it fetches the 16-byte aligned span around the (8-byte aligned) row.  Slot
// strides are ODD so that the 32 points of a warp reading the same wire fall into different bank pairs.
struct TileGeom {
  u32 tw;                       // point warps per tile
  u32 ws, css, zss;             // slot strides of the three staged matrices (odd, >= words + 3)
  u32 nw, ncs, nzs;             // words copied per row
  __host__ __device__ u32 points() const { return tw * 32; }
  __host__ __device__ u32 groups() const { return QUOT_WARPS / tw; }
  __host__ __device__ size_t words() const {
    return 2 + 16 /* queue */ + (size_t)points() * (ws + css + zss + MAX_CHALLENGES) + (size_t)groups() * MAX_CHALLENGES * points();
  }
  __host__ static u32 odd_stride(u32 words) { return (words + 3) | 1; }
};

// cp.async.bulk global -> shared of `bytes` (multiple of 16) from a 16-byte aligned source, completing on `mbar`
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, u32 bytes, u64* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (u32)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((u32)__cvta_generic_to_shared(mbar))
               : "memory");
}

// Stages one row (nwords u64 at src, 8-byte aligned) into the slot at `slot`: the bulk copy fetches the enclosing 16-byte
// aligned span [src - lead, ...) into the slot's first 16-byte aligned word; *row_out points at the row inside the slot.
__device__ __forceinline__ void stage_row(u64* slot, const u64* src, u32 nwords, u64* mbar, u64** row_out) {
  u64* dst = slot + ((((u64)__cvta_generic_to_shared(slot)) >> 3) & 1);   // 16-byte aligned
  const u32 lead = (u32)((((u64)src) >> 3) & 1);                          // 1 if the row starts in the upper half of a 16-byte unit
  const u32 bytes = ((lead + nwords + 1) & ~1u) * 8;
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
  bulk_load(dst, src - lead, bytes, mbar);
  *row_out = dst + lead;
}

template <int NC>
__global__ void __launch_bounds__(QUOT_WARPS * 32) quotient_values_kernel(Params p, TileGeom tg, const u32* __restrict__ work_items,
                                                                          u32 num_items) {
  extern __shared__ __align__(16) u64 qsh[];
  const u32 TP = tg.points(), G = tg.groups();
  // layout: [mbar (2 words)] [queue: counters (2 x u32), per-group item slots (2 x 12 x u32)] [wires TP x ws] [cs TP x css]
  //         [zs TP x zss] [zn TP x MAX_CH] [part G x NC x TP]
  u64* mbar = qsh;
  u32* queue = reinterpret_cast<u32*>(qsh + 2);   // [0], [1]: next item (optimistic pass, exact pass); [2 + 2g + parity]: group g's item
  u64* sw = qsh + 2 + 16;
  u64* scs = sw + (size_t)TP * tg.ws;
  u64* szs = scs + (size_t)TP * tg.css;
  u64* szn = szs + (size_t)TP * tg.zss;
  u64* part = szn + (size_t)TP * MAX_CHALLENGES;
  __shared__ const u64* row_w[QUOT_WARPS * 32 / 2];   // >= TP (TP <= 192)
  __shared__ const u64* row_cs[QUOT_WARPS * 32 / 2];
  __shared__ const u64* row_zs[QUOT_WARPS * 32 / 2];
  const u64 lde_size = (u64)1 << (p.degree_bits + p.qdb);
  const u64 tile_first = (u64)blockIdx.x * TP;   // in units of this launch's point list (j)
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 lde_bits = p.degree_bits + p.rate_bits, step_log = p.rate_bits - p.qdb;

  // ---- phase 1: stage the rows ----
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    queue[0] = 0;
    queue[1] = 0;
  }
  __syncthreads();
  if (tid < TP) {
    u64 j = tile_first + tid;
    if (j >= p.pt_count) j = p.pt_count - 1;  // tail slots recompute the last point; their results are not stored
    const u64 i = p.pt_first + p.pt_stride * j;
    const u64 row = lde_bits ? (__brevll(i << step_log) >> (64 - lde_bits)) : 0;
    const u64 i_next = (i + ((u64)1 << p.qdb)) & (lde_size - 1);
    const u64 row_next = lde_bits ? (__brevll(i_next << step_log) >> (64 - lde_bits)) : 0;
    u64 *rw, *rc, *rz;
    // the last row of a matrix may not be followed by a readable word: it is copied with ordinary loads
    const bool last = row == p.last_row;
    if (!last) {
      stage_row(sw + (size_t)tid * tg.ws, p.wires + row * p.wires_stride, tg.nw, mbar, &rw);
      stage_row(scs + (size_t)tid * tg.css, p.cs + row * p.cs_stride, tg.ncs, mbar, &rc);
      stage_row(szs + (size_t)tid * tg.zss, p.zs_pp + row * p.zs_stride, tg.nzs, mbar, &rz);
    } else {
      rw = sw + (size_t)tid * tg.ws;
      rc = scs + (size_t)tid * tg.css;
      rz = szs + (size_t)tid * tg.zss;
      for (u32 k = 0; k < tg.nw; k++) rw[k] = p.wires[row * p.wires_stride + k];
      for (u32 k = 0; k < tg.ncs; k++) rc[k] = p.cs[row * p.cs_stride + k];
      for (u32 k = 0; k < tg.nzs; k++) rz[k] = p.zs_pp[row * p.zs_stride + k];
    }
    row_w[tid] = rw;
    row_cs[tid] = rc;
    row_zs[tid] = rz;
    for (u32 c = 0; c < (u32)NC; c++) szn[(size_t)tid * MAX_CHALLENGES + c] = __ldg(p.zs_pp + row_next * p.zs_stride + c);
  }
  __syncthreads();   // all expect_tx are registered before the single arrival
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((u32)__cvta_generic_to_shared(mbar)) : "memory");
  {
    u32 done = 0;
    while (!done)
      asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                   : "=r"(done)
                   : "r"((u32)__cvta_generic_to_shared(mbar)), "r"(0)
                   : "memory");
  }
  __syncthreads();

  // ---- phase 2: warp (group, point warp) evaluates its group's work items for its 32 points ----
  const u32 group = wid / tg.tw, slot = (wid % tg.tw) * 32 + lane;
  const u64 i = p.pt_first + p.pt_stride * min(tile_first + slot, p.pt_count - 1);
  PointRows r;
  r.w = row_w[slot];
  r.cs = row_cs[slot];
  r.zp = row_zs[slot];
  r.zn = szn + (size_t)slot * MAX_CHALLENGES;
  r.x = 0;
  const u32 pw = wid % tg.tw;
  auto run = [&](auto& mode, u32* counter) {
    u64 acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0;
#pragma unroll 1
    for (u32 it = 0;; it++) {
      // the group's first warp takes the next item off the queue; a named barrier (one per group) publishes it
      u32* slot_item = queue + 2 + 2 * group + (it & 1);
      if (pw == 0 && lane == 0) *slot_item = atomicAdd(counter, 1u);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(tg.tw * 32) : "memory");
      const u32 idx = *slot_item;
      if (idx >= num_items) break;
      const u32 item = work_items[idx];
      u64 t[NC];
      if (item == WORK_PERMUTATION) {
        eval_permutation_terms(p, r, i, mode, t);
      } else {
        const GateDesc g = p.gates[item];
        eval_gate_terms(p, g, item, r, mode, t);
      }
#pragma unroll
      for (int c = 0; c < NC; c++) acc[c] = gl::add(acc[c], t[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) part[((size_t)group * NC + c) * TP + slot] = acc[c];
  };
#ifndef P2B_EXACT_ONLY
  gl::Optimistic fast;
  run(fast, queue);
  // an optimistic reduction hit its rare case somewhere in this CTA's points: redo the tile exactly
  if (__syncthreads_or(fast.rare))
#endif
  {
    gl::Exact exact;
    run(exact, queue + 1);
  }
  __syncthreads();
  // ---- phase 3: sum the groups, divide by Z_H (prover.rs:985-991) ----
  if (tid < TP && tile_first + tid < p.pt_count) {
    const u64 pj = tile_first + tid, pi = p.pt_first + p.pt_stride * pj;
    const u64 zi = p.zh_inv[pi & ((1u << p.qdb) - 1)];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      u64 v = 0;
      for (u32 g2 = 0; g2 < G; g2++) v = gl::add(v, part[((size_t)g2 * NC + c) * TP + tid]);
      v = gl::canon(gl::mul(v, zi));
      p.out_values[(u64)c * p.pt_count + pj] = v;
      if (p.out_rows) p.out_rows[pj * NC + c] = v;
    }
  }
}

// host: rough thread-instruction cost of one work item (only the RATIOS matter: they balance the groups)
inline u64 work_item_cost(const Params& p, const GateDesc* g) {
  const u64 nc = p.num_challenges, emit = 12 * nc + 6;
  if (!g) return (u64)p.num_routed * nc * 70 + (u64)(p.num_partial_products + 2) * nc * (emit + 40) + 1500;
  const u64 k = gate_num_constraints(*g);
  u64 extra = 0;
  switch (g->type) {
    case G_ARITHMETIC: extra = (u64)g->p0 * 50; break;
    case G_BASE_SUM: extra = (u64)g->p0 * (20 + 14 * g->p1); break;
    case G_POSEIDON: extra = 17000; break;
    case G_RANDOM_ACCESS: extra = (u64)g->p1 * ((1u << g->p0) * 25 + g->p0 * 30); break;
    case G_U32_ARITHMETIC: extra = (u64)g->p0 * 36 * 45; break;
    case G_U32_ADD_MANY: extra = (u64)g->p1 * 21 * 45; break;
    case G_U32_RANGE_CHECK: extra = (u64)g->p0 * 17 * 45; break;
    case G_U32_SUBTRACTION: extra = (u64)g->p0 * 19 * 45; break;
    case G_COMPARISON: extra = k * 60; break;
    case G_ARITHMETIC_EXT: extra = (u64)g->p0 * 130; break;
    case G_MUL_EXT: extra = (u64)g->p0 * 90; break;
    case G_REDUCING: extra = (u64)g->p0 * 70; break;
    case G_REDUCING_EXT: extra = (u64)g->p0 * 110; break;
    case G_EXPONENTIATION: extra = (u64)g->p0 * 45; break;
    case G_POSEIDON_MDS: extra = 9000; break;
    case G_HIGH_DEGREE_INTERPOLATION: extra = (u64)(1u << g->p0) * 400; break;
    case G_LOW_DEGREE_INTERPOLATION: extra = (u64)(1u << g->p0) * 500; break;
    default: break;
  }
  return k * emit + extra + 200;
}

// multi-device: values[c][i] <- parts[(src(i) * nc + c) * count + i / G] where device src(i) = reverse_bits(i mod G) evaluated
// the points congruent to i mod G (csrc/mgpu.cuh)
__global__ void interleave_parts_kernel(const u64* __restrict__ parts, u64* __restrict__ values, u64 lde_size, u32 nc, u32 g_log) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lde_size) return;
  const u64 G = (u64)1 << g_log, count = lde_size >> g_log;
  const u64 res = i & (G - 1);
  const u64 src = g_log ? (__brevll(res) >> (64 - g_log)) : 0;
  for (u32 c = 0; c < nc; c++) values[(u64)c * lde_size + i] = parts[(src * nc + c) * count + (i >> g_log)];
}

// coefficients[i] *= shift_inv^i  (coset_ifft, field/src/polynomial/mod.rs:64-77)
__global__ void scale_by_powers_kernel(u64* __restrict__ v, u64 len, u64 ncols, u64 base) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  u64 pw = gl::pow(base, i);
  for (u64 c = 0; c < ncols; c++) v[c * len + i] = gl::canon(gl::mul(v[c * len + i], pw));
}

// alpha_pows[c][t] = alpha_c^t
__global__ void alpha_pows_kernel(u64* __restrict__ out, u32 num_terms, u32 nc, const u64* __restrict__ alphas) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_terms) return;
  for (u32 c = 0; c < nc; c++) out[(u64)c * num_terms + t] = gl::canon(gl::pow(alphas[c], t));
}

}  // namespace quotient
