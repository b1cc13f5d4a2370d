// gl64.cuh -- Goldilocks field (p = 2^64 - 2^32 + 1) for sm_100a integer pipes.
//
// Semantics follow the reference's GoldilocksField (field/src/goldilocks_field.rs): an element is ANY u64
// (values in [p, 2^64) are legal non-canonical representatives, :169-176), arithmetic results are u64
// representatives of the exact field result, and everything that leaves the device is canonicalised
// (gl_canon) because the reference compares / serialises canonical values (:34-38).
//
// Instruction budget (checked with cuobjdump -sass): gl_mul = 4 IMAD.WIDE.U32 (fma pipe) for the 128-bit
// product + 1 IMAD.WIDE.U32 for hi_lo*eps + ~10 IADD3/SEL (alu pipe).  There is no 64-bit integer ALU on
// sm_100a; every u64 add is an IADD3 + IADD3.X pair, so carries are kept in predicate registers through
// add.cc/addc chains instead of being recomputed with compares.
#pragma once
#include <stdint.h>

namespace gl {

typedef uint64_t u64;
typedef uint32_t u32;

static constexpr u64 P = 0xFFFFFFFF00000001ull;
static constexpr u64 EPS = 0xFFFFFFFFull;

__device__ __forceinline__ void split(u64 x, u32& lo, u32& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
__device__ __forceinline__ u64 pack(u32 lo, u32 hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}

// to_canonical_u64 (goldilocks_field.rs:169-176)
__device__ __forceinline__ u64 canon(u64 x) { return x >= P ? x - P : x; }

// a + b for arbitrary u64 representatives (goldilocks_field.rs:197-219).  2^64 == eps (mod p), so each
// wrap of the 64-bit adder is repaid with +eps; the second wrap can only happen when both inputs were
// non-canonical.
__device__ __forceinline__ u64 add(u64 a, u64 b) {
  u64 s;
  u32 c;
  asm("{ add.cc.u64 %0, %2, %3; addc.u32 %1, 0, 0; }" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  u64 adj = (u64)(0u - c);  // 0 or eps
  asm("{ add.cc.u64 %0, %0, %2; addc.u32 %1, 0, 0; }" : "+l"(s), "=r"(c) : "l"(adj));
  return s + (u64)(0u - c);
}
// a + b where b is canonical (< p): a single wrap repayment suffices (add_canonical_u64, :154-158).
__device__ __forceinline__ u64 add_canonical(u64 a, u64 b) {
  u64 s;
  u32 c;
  asm("{ add.cc.u64 %0, %2, %3; addc.u32 %1, 0, 0; }" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  return s + (u64)(0u - c);
}
// a - b for arbitrary representatives (goldilocks_field.rs:234-256)
__device__ __forceinline__ u64 sub(u64 a, u64 b) {
  u64 d;
  u32 w;
  asm("{ sub.cc.u64 %0, %2, %3; subc.u32 %1, 0, 0; }" : "=l"(d), "=r"(w) : "l"(a), "l"(b));  // w = -borrow
  u64 adj = (u64)w;  // 0 or eps
  asm("{ sub.cc.u64 %0, %0, %2; subc.u32 %1, 0, 0; }" : "+l"(d), "=r"(w) : "l"(adj));
  return d - (u64)w;
}
// a - b where b is canonical: one repayment (sub_canonical_u64, :161-165)
__device__ __forceinline__ u64 sub_canonical(u64 a, u64 b) {
  u64 d;
  u32 w;
  asm("{ sub.cc.u64 %0, %2, %3; subc.u32 %1, 0, 0; }" : "=l"(d), "=r"(w) : "l"(a), "l"(b));
  return d - (u64)w;
}
__device__ __forceinline__ u64 neg(u64 a) {
  u64 c = canon(a);
  return c ? P - c : 0;
}

// 64x64 -> 128 (ptxas CSEs the two halves into 4 IMAD.WIDE.U32 + 3 fix-ups)
__device__ __forceinline__ void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
  asm("{ mul.lo.u64 %0, %2, %3; mul.hi.u64 %1, %2, %3; }" : "=&l"(lo), "=&l"(hi) : "l"(a), "l"(b));
}

// reduce128 (goldilocks_field.rs:345-358): x = lo + 2^64*hi,  2^64 == eps,  2^96 == -1:
//   x == lo + hi_lo*eps - hi_hi.
// Order used here: y = lo + hi_lo*eps (one IMAD.WIDE with carry-out c), r = y - hi_hi (borrow b).
//   c = 1: true value is y + 2^64 - hi_hi == y + eps - hi_hi, which fits u64 exactly (y < hi_lo*eps), so
//          r + eps (mod 2^64) is the answer whether or not the subtraction borrowed.
//   c = 0, b = 1: true value is negative; r wrapped by +2^64, repay with -eps (cannot underflow).
__device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
  u32 h0, h1;
  split(hi, h0, h1);
  u64 y, r;
  u32 c, w;
  asm("{ .reg .u64 m; mul.wide.u32 m, %2, 0xffffffff; add.cc.u64 %0, m, %3; addc.u32 %1, 0, 0; }"
      : "=l"(y), "=r"(c)
      : "r"(h0), "l"(lo));
  asm("{ sub.cc.u64 %0, %2, %3; subc.u32 %1, 0, 0; }" : "=l"(r), "=r"(w) : "l"(y), "l"((u64)h1));  // w = -b
  r = (u64)c * EPS + r;            // arithmetic mod 2^64 is intended
  u32 k = w & (c - 1u);            // 0xffffffff iff (b && !c)
  return r - (u64)k;               // - eps
}

__device__ __forceinline__ u64 mul(u64 a, u64 b) {
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi);
}

// ---- "optimistic" reduction ---------------------------------------------------------------------------------
// In reduce128 the repayment for (borrow && !carry) costs 5 of the 10 instructions but fires only when
// lo + hi_lo*eps < hi_hi < 2^32, i.e. with probability ~2^-32 per multiplication on random data.  The optimistic form
// leaves that case out and records it in `rare` (one PLOP3 on the two carry predicates); a caller that finds `rare` set
// discards its result and recomputes with the exact functions above, so results stay bit-exact for EVERY input (the
// tests force the case with x = 2^48: x*x = 2^96 has lo = 0, hi = 2^32).
struct Exact {
  static constexpr bool optimistic = false;
  bool rare = false;
  u32 lazy_add = 0, lazy_sub = 0;   // never written in this mode (the templates below compile for both)
  __device__ __forceinline__ bool any() const { return false; }
};
struct Optimistic {
  static constexpr bool optimistic = true;
  bool rare = false;
  u32 lazy_add = 0, lazy_sub = 0;   // second wraps of the optimistic add (counted up) / sub (counted down) below
  __device__ __forceinline__ bool any() const { return rare | ((lazy_add | lazy_sub) != 0); }
};

template <class M>
__device__ __forceinline__ u64 reduce128(u64 lo, u64 hi, M& m) {
  if (!M::optimistic) return reduce128(lo, hi);
  u32 h0, h1;
  split(hi, h0, h1);
  u64 y, r;
  u32 c, w;
  asm("{ .reg .u64 m; mul.wide.u32 m, %2, 0xffffffff; add.cc.u64 %0, m, %3; addc.u32 %1, 0, 0; }"
      : "=l"(y), "=r"(c)
      : "r"(h0), "l"(lo));
  asm("{ sub.cc.u64 %0, %2, %3; subc.u32 %1, 0, 0; }" : "=l"(r), "=r"(w) : "l"(y), "l"((u64)h1));
  m.rare |= (w != 0) & (c == 0);
  return (u64)c * EPS + r;
}
// Optimistic add / sub for arbitrary u64 representatives: ONE repayment of the 64-bit wrap (2^64 == eps) instead of two.
// The second wrap needs a + b - 2^64 >= p (both operands within 2^32 of 2^64; probability ~2^-64 on random data, but legal
// inputs can force it), so its carry / borrow is summed into m.lazy_add / m.lazy_sub on the carry chain itself -- 7 and 6
// SASS instructions against 10 for the exact forms; the caller redoes its work exactly when m.any().  Each chain stays of
// one kind (add.cc -> addc, sub.cc -> subc): the carry flag read by the other kind is not what the PTX text suggests.
template <class M>
__device__ __forceinline__ u64 add(u64 a, u64 b, M& m) {
  if (!M::optimistic) return add(a, b);
  u64 s;
  u32 c;
  asm("{ add.cc.u64 %0, %2, %3; addc.u32 %1, 0, 0; }" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  u64 adj = (u64)(0u - c);  // 0 or eps
  asm("{ add.cc.u64 %0, %0, %2; addc.u32 %1, %1, 0; }" : "+l"(s), "+r"(m.lazy_add) : "l"(adj));
  return s;
}
template <class M>
__device__ __forceinline__ u64 sub(u64 a, u64 b, M& m) {
  if (!M::optimistic) return sub(a, b);
  u64 d;
  u32 w;
  asm("{ sub.cc.u64 %0, %2, %3; subc.u32 %1, 0, 0; }" : "=l"(d), "=r"(w) : "l"(a), "l"(b));  // w = -borrow: 0 or eps
  asm("{ sub.cc.u64 %0, %0, %2; subc.u32 %1, %1, 0; }" : "+l"(d), "+r"(m.lazy_sub) : "l"((u64)w));   // counts down
  return d;
}
template <class M>
__device__ __forceinline__ u64 mul(u64 a, u64 b, M& m) {
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi, m);
}
template <class M>
__device__ __forceinline__ u64 sqr(u64 a, M& m) { return mul(a, a, m); }
template <class M>
__device__ __forceinline__ u64 mul_add(u64 a, u64 b, u64 c, M& m) {
  u64 lo, hi;
  asm("{ mad.lo.cc.u64 %0, %2, %3, %4; madc.hi.u64 %1, %2, %3, 0; }" : "=&l"(lo), "=&l"(hi) : "l"(a), "l"(b), "l"(c));
  return reduce128(lo, hi, m);
}
__device__ __forceinline__ u64 sqr(u64 a) { return mul(a, a); }

// a*b + c (multiply_accumulate, goldilocks_field.rs:123-127): u64 + u64*u64 cannot overflow 128 bits
__device__ __forceinline__ u64 mul_add(u64 a, u64 b, u64 c) {
  u64 lo, hi;
  asm("{ mad.lo.cc.u64 %0, %2, %3, %4; madc.hi.u64 %1, %2, %3, 0; }" : "=&l"(lo), "=&l"(hi) : "l"(a), "l"(b), "l"(c));
  return reduce128(lo, hi);
}

// from_noncanonical_u96 (goldilocks_field.rs:153-165): lo + 2^64*hi32
__device__ __forceinline__ u64 reduce96(u64 lo, u32 hi) {
  u64 y;
  u32 c;
  asm("{ .reg .u64 m; mul.wide.u32 m, %2, 0xffffffff; add.cc.u64 %0, m, %3; addc.u32 %1, 0, 0; }"
      : "=l"(y), "=r"(c)
      : "r"(hi), "l"(lo));
  return (u64)c * EPS + y;
}

// x^e by square-and-multiply (exp_u64, field/src/types.rs:347-371)
__device__ __forceinline__ u64 pow(u64 base, u64 e) {
  u64 cur = base, acc = 1;
  while (e) {
    if (e & 1) acc = mul(acc, cur);
    cur = sqr(cur);
    e >>= 1;
  }
  return acc;
}

// inverse_2exp(e), e <= 32 (field/src/types.rs:227-266)
__host__ __device__ __forceinline__ u64 inverse_2exp(unsigned e) { return P - ((P - 1) >> e); }

}  // namespace gl
