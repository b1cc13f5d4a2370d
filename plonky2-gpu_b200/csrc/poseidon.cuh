// poseidon.cuh -- Poseidon permutation over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds) and the
// overwrite-mode sponge / 2-to-1 compression built on it, for sm_100a.
//
// What is computed (bit-exact with the reference CPU path):
//   permute()        Poseidon::poseidon            plonky2/src/hash/poseidon.rs:590-606
//   full round       constant_layer/sbox/mds_layer poseidon.rs:482-493, 525-548, 172-260
//   partial rounds   "fast" form                   poseidon.rs:574-588 with :310-365 (first-round constants +
//                                                  11x11 initial matrix) and :398-427 (W_HATS / VS layer)
//   hash_or_noop / hash_no_pad / two_to_one        plonk/config.rs:56-67, hash/hashing.rs:81-104, :65-72
//
// How (B200): the 12-word state lives in registers (24 x 32-bit); a full round's MDS works on the 32-bit
// halves of each word with IMAD.WIDE.U32 multiply-accumulates by the <=6-bit circulant entries (immediates
// in the instruction stream), sums both halves in 64-bit accumulators (no carries possible: 12 * 41 * 2^32
// < 2^42), adds the NEXT round's constant into the 96-bit sum and reduces once.  S-boxes are 4 mod-muls of
// gl::mul (4 IMAD.WIDE + reduce).  Round loops are kept rolled (#pragma unroll 1) so the hot loop bodies
// (~1.2k instructions for a full round, ~0.5k for a partial round) stay inside the instruction cache;
// per-round constants are fetched from __constant__ memory (uniform across the warp -> LDC broadcast).
#pragma once
#include "gl64.cuh"
#include "poseidon_tables.h"

namespace poseidon {

using gl::u32;
using gl::u64;

// __constant__ copies (filled by poseidon_upload_constants()).
struct Consts {
  u64 rc[360];        // ALL_ROUND_CONSTANTS
  u64 first_rc[12];   // FAST_PARTIAL_FIRST_ROUND_CONSTANT
  u64 partial_rc[22]; // FAST_PARTIAL_ROUND_CONSTANTS
  u64 vs[22 * 11];    // FAST_PARTIAL_ROUND_VS
  u64 w_hats[22 * 11];// FAST_PARTIAL_ROUND_W_HATS
  u64 init[11 * 11];  // FAST_PARTIAL_ROUND_INITIAL_MATRIX  [r-1][c-1]
  u64 post[8 * 12];   // constants added right after the MDS of full round k (k = 0..7): rows 1,2,3 of rc,
                      // first_rc, rows 27,28,29 of rc, zeros  (next round's constant_layer folded forward)
};
// The library is a single translation unit (plonky2_b200.cu), so the definition lives here.
__constant__ Consts C;

// MDS circulant first row and diagonal (poseidon_goldilocks.rs:21-22) as compile-time immediates.
__device__ __forceinline__ constexpr u32 mds_circ(int i) {
  constexpr u32 t[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  return t[i];
}
static constexpr u32 MDS_DIAG0 = 8;

// x^7 (sbox_monomial, poseidon.rs:525-532)
template <class M>
__device__ __forceinline__ u64 sbox(u64 x, M& m) {
  u64 x2 = gl::sqr(x, m);
  u64 x4 = gl::sqr(x2, m);
  u64 x3 = gl::mul(x, x2, m);
  return gl::mul(x3, x4, m);
}
__device__ __forceinline__ u64 sbox(u64 x) {
  gl::Exact m;
  return sbox(x, m);
}

// out[r] = sum_i circ[i] * s[(i+r)%12] + diag[r]*s[r] + addc[r]   (mds_row_shf + mds_layer, then the next
// constant_layer folded in).  addc must be canonical round constants (< p).
__device__ __forceinline__ void mds_layer(u64 (&s)[12], const u64* __restrict__ addc) {
#ifdef P2B_MDS_F64
  // FP64-pipe form: the 32-bit halves of every word as exact doubles, 12 DFMA per output half accumulated onto
  // 2^52 + (half of the round constant), so the integer sum sits in the mantissa.
  double lo[12], hi[12];
  const double M52 = 4503599627370496.0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u32 l, h;
    gl::split(s[i], l, h);
    lo[i] = __hiloint2double(0x43300000, (int)l) - M52;
    hi[i] = __hiloint2double(0x43300000, (int)h) - M52;
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u32 c0, c1;
    gl::split(addc[r], c0, c1);
    double al = __hiloint2double(0x43300000, (int)c0), ah = __hiloint2double(0x43300000, (int)c1);
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al = fma(lo[(i + r) % 12], (double)mds_circ(i), al);
      ah = fma(hi[(i + r) % 12], (double)mds_circ(i), ah);
    }
    if (r == 0) {
      al = fma(lo[0], (double)MDS_DIAG0, al);
      ah = fma(hi[0], (double)MDS_DIAG0, ah);
    }
    u32 w0 = (u32)__double2loint(al), l1 = (u32)__double2hiint(al), h0 = (u32)__double2loint(ah), h1 = (u32)__double2hiint(ah), w1, w2;
    asm("{ .reg .u32 t; sub.u32 t, %2, 0x43300000; add.cc.u32 %0, t, %3; .reg .u32 u; sub.u32 u, %4, 0x43300000; addc.u32 %1, u, 0; }"
        : "=r"(w1), "=r"(w2) : "r"(l1), "r"(h0), "r"(h1));
    s[r] = gl::reduce96(gl::pack(w0, w1), w2);
  }
#elif defined(P2B_MDS_LIMB22)
  // Three 22/22/20-bit limbs per word: every product c*limb and every 12-term sum stays below 2^32, so the whole
  // layer is 32-bit IMAD (64 thread-instr/clk/SM) instead of IMAD.WIDE (~23): 432 IMAD vs 288 IMAD.WIDE.
  u32 x0[12], x1[12], x2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u32 lo, hi;
    gl::split(s[i], lo, hi);
    x0[i] = lo & 0x3fffffu;
    x1[i] = __funnelshift_r(lo, hi, 22) & 0x3fffffu;
    x2[i] = hi >> 12;
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u32 s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      s0 += x0[(i + r) % 12] * mds_circ(i);
      s1 += x1[(i + r) % 12] * mds_circ(i);
      s2 += x2[(i + r) % 12] * mds_circ(i);
    }
    if (r == 0) {
      s0 += x0[0] * MDS_DIAG0;
      s1 += x1[0] * MDS_DIAG0;
      s2 += x2[0] * MDS_DIAG0;
    }
    // value = s0 + s1*2^22 + s2*2^44 + addc[r]  ->  (w0, w1, w2)
    u32 c0, c1, w0, w1, w2;
    gl::split(addc[r], c0, c1);
    u32 a_lo = s1 << 22, a_hi = s1 >> 10;   // s1 * 2^22 as (lo, hi)
    u32 b_lo = s2 << 12, b_hi = s2 >> 20;   // s2 * 2^44 as (mid, top)
    asm("{\n\t"
        "add.cc.u32 %0, %3, %4;\n\t"        // w0 = s0 + a_lo
        "addc.cc.u32 %1, %5, %6;\n\t"       // w1 = a_hi + b_lo + c
        "addc.u32 %2, %7, 0;\n\t"           // w2 = b_hi + c
        "add.cc.u32 %0, %0, %8;\n\t"        // + round constant
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.u32 %2, %2, 0;\n\t"
        "}"
        : "=&r"(w0), "=&r"(w1), "=&r"(w2)
        : "r"(s0), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(c0), "r"(c1));
    s[r] = gl::reduce96(gl::pack(w0, w1), w2);
  }
#else
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) gl::split(s[i], lo[i], hi[i]);
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u64 al = 0, ah = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += (u64)lo[(i + r) % 12] * mds_circ(i);
      ah += (u64)hi[(i + r) % 12] * mds_circ(i);
    }
    if (r == 0) {
      al += (u64)lo[0] * MDS_DIAG0;
      ah += (u64)hi[0] * MDS_DIAG0;
    }
    // value = al + ah * 2^32 (+ addc)  as (w0, w1, w2) 32-bit limbs -> 96 bits
    u32 al0, al1, ah0, ah1, w0, w1, w2;
    gl::split(al, al0, al1);
    gl::split(ah, ah0, ah1);
    {
      u32 c0, c1;
      gl::split(addc[r], c0, c1);
      asm("{\n\t"
          "add.cc.u32 %1, %4, %5;\n\t"
          "addc.u32 %2, %6, 0;\n\t"
          "add.cc.u32 %0, %3, %7;\n\t"
          "addc.cc.u32 %1, %1, %8;\n\t"
          "addc.u32 %2, %2, 0;\n\t"
          "}"
          : "=&r"(w0), "=&r"(w1), "=&r"(w2)
          : "r"(al0), "r"(al1), "r"(ah0), "r"(ah1), "r"(c0), "r"(c1));
    }
    s[r] = gl::reduce96(gl::pack(w0, w1), w2);
  }
#endif
}

// sbox_layer (poseidon.rs:534-548).  Code size matters more than instruction count here: with ~28 warps per SM
// streaming through the permutation, instruction fetch stalls ("no_instruction" in ncu) dominate as soon as the
// hot code exceeds the ~32 KB instruction cache, so the 12 S-boxes are a rolled loop of 3 x 4 with a register
// rotation (24 moves per iteration) instead of 12 inlined copies (13 KB of SASS).
template <class M>
__device__ __forceinline__ void sbox_layer(u64 (&s)[12], M& m) {
#ifdef P2B_SBOX_UNROLLED
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = sbox(s[i], m);
#else
#pragma unroll 1
  for (int g = 0; g < 3; g++) {
    u64 t0 = sbox(s[0], m), t1 = sbox(s[1], m), t2 = sbox(s[2], m), t3 = sbox(s[3], m);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = s[i + 4];
    s[8] = t0;
    s[9] = t1;
    s[10] = t2;
    s[11] = t3;
  }
#endif
}

__device__ __forceinline__ void sbox_layer(u64 (&s)[12]) {
  gl::Exact m;
  sbox_layer(s, m);
}

// 128-bit accumulate helper for the partial-round dot products: (acc_lo, acc_hi, acc_top) += a*b
__device__ __forceinline__ void mac160(u64& lo, u64& hi, u32& top, u64 a, u64 b) {
  u64 pl, ph;
  gl::mul_wide(a, b, pl, ph);
  asm("{ add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u32 %2, %2, 0; }" : "+l"(lo), "+l"(hi), "+r"(top) : "l"(pl), "l"(ph));
}
// reduce_u160 (poseidon.rs:40-47)
template <class M>
__device__ __forceinline__ u64 reduce160(u64 lo, u64 hi, u32 top, M& m) {
  u64 reduced_hi = gl::reduce96(hi, top);
  return gl::reduce128(lo, reduced_hi, m);
}
__device__ __forceinline__ u64 reduce160(u64 lo, u64 hi, u32 top) {
  gl::Exact m;
  return reduce160(lo, hi, top, m);
}

// Optional block-wide barrier at every round boundary (P2B_SYNC_ROUNDS): keeps all warps of a CTA inside the same
// loop body so they share instruction-cache lines (the unrolled round bodies are 8-20 KB each and ncu shows
// "no_instruction" as the top stall when warps drift apart).  Only legal when every thread of the CTA runs the
// same number of permutations -- the kernels that enable it keep their tail threads alive on clamped indices.
__device__ __forceinline__ void round_sync() {
#ifdef P2B_SYNC_ROUNDS
  __syncthreads();
#endif
}

// mds_partial_layer_init (poseidon.rs:310-337): s[0] unchanged; s[c] = sum_r s[r] * init[r-1][c-1].
// Rolled over the output column c (keeps ~24 KB of straight-line code out of the instruction cache); the results
// are pushed through a shift register so no register array is indexed dynamically.
template <class M>
__device__ __forceinline__ void partial_layer_init(u64 (&s)[12], M& m) {
#ifdef P2B_INIT_UNROLLED
  u64 t[12];
  t[0] = s[0];
#pragma unroll
  for (int c = 1; c < 12; c++) {
    u64 lo = 0, hi = 0;
    u32 top = 0;
#pragma unroll
    for (int r = 1; r < 12; r++) mac160(lo, hi, top, s[r], C.init[(r - 1) * 11 + (c - 1)]);
    t[c] = reduce160(lo, hi, top, m);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = t[i];
#else
  u64 t[12];
#pragma unroll
  for (int i = 1; i < 12; i++) t[i] = 0;
#pragma unroll 1
  for (int c = 0; c < 11; c++) {
    u64 lo = 0, hi = 0;
    u32 top = 0;
#pragma unroll
    for (int r = 1; r < 12; r++) mac160(lo, hi, top, s[r], C.init[(r - 1) * 11 + c]);
    u64 v = reduce160(lo, hi, top, m);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = t[i + 1];
    t[11] = v;
  }
#pragma unroll
  for (int i = 1; i < 12; i++) s[i] = t[i];
#endif
}

// mds_partial_layer_fast (poseidon.rs:398-427) for partial round r; s0 = the S-boxed (and constant-added) lane 0:
//   s[0] <- 25*s0 + sum_i w_hat[r][i-1] * s[i]   (u160 accumulator),   s[i] <- s[i] + s0 * vs[r][i-1]
template <class M>
__device__ __forceinline__ void partial_layer_fast(u64 (&s)[12], u64 s0, int r, M& m) {
  u64 lo, hi;
  u32 top = 0;
  {
    u32 a0, a1;
    gl::split(s0, a0, a1);
    u64 pl = (u64)a0 * (mds_circ(0) + MDS_DIAG0), ph = (u64)a1 * (mds_circ(0) + MDS_DIAG0);
    u32 pl0, pl1, ph0, ph1, m1, m2;
    gl::split(pl, pl0, pl1);
    gl::split(ph, ph0, ph1);
    asm("{ add.cc.u32 %0, %2, %3; addc.u32 %1, %4, 0; }" : "=&r"(m1), "=&r"(m2) : "r"(pl1), "r"(ph0), "r"(ph1));
    lo = gl::pack(pl0, m1);
    hi = (u64)m2;
  }
#pragma unroll
  for (int i = 1; i < 12; i++) mac160(lo, hi, top, s[i], C.w_hats[r * 11 + i - 1]);
  u64 d = reduce160(lo, hi, top, m);
#pragma unroll
  for (int i = 1; i < 12; i++) s[i] = gl::mul_add(s0, C.vs[r * 11 + i - 1], s[i], m);
  s[0] = d;
}
__device__ __forceinline__ void partial_layer_init(u64 (&s)[12]) {
  gl::Exact m;
  partial_layer_init(s, m);
}
__device__ __forceinline__ void partial_layer_fast(u64 (&s)[12], u64 s0, int r) {
  gl::Exact m;
  partial_layer_fast(s, s0, r, m);
}

// The permutation.  Input: any u64 representatives; output: u64 representatives (NOT canonicalised --
// callers canonicalise what they store).
template <class M>
__device__ __forceinline__ void permute(u64 (&s)[12], M& m) {
  // constant layer of round 0 up front; every later constant layer is folded into the preceding MDS.
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_canonical(s[i], C.rc[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    // ---- 4 full rounds (poseidon.rs:560-572) ----
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
      round_sync();
      sbox_layer(s, m);
      mds_layer(s, &C.post[12 * (half * 4 + r)]);
    }
    if (half == 0) {
      // ---- partial rounds (poseidon.rs:574-588); first-round constants were folded into post[3] ----
      partial_layer_init(s, m);
#pragma unroll 1
      for (int r = 0; r < 22; r++) {
        round_sync();
        u64 s0 = gl::add_canonical(sbox(s[0], m), C.partial_rc[r]);
        partial_layer_fast(s, s0, r, m);
      }
      // constant layer of round 26 (first of the closing full rounds)
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = gl::add_canonical(s[i], C.rc[12 * 26 + i]);
    }
  }
}

// exact permutation (every reduction fully repaid)
__device__ __forceinline__ void permute(u64 (&s)[12]) {
  gl::Exact m;
  permute(s, m);
}

// compress / two_to_one (hashing.rs:65-72): perm([l, r, 0,0,0,0])[0..4]; output canonical.
template <class M>
__device__ __forceinline__ void two_to_one(const u64 l[4], const u64 r[4], u64 out[4], M& m) {
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s[i] = l[i];
    s[4 + i] = r[i];
    s[8 + i] = 0;
  }
  permute(s, m);
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
}

// host: copy the tables into __constant__ memory of the current device
inline cudaError_t upload_constants() {
  static Consts h;
  for (int i = 0; i < 360; i++) h.rc[i] = P2_ROUND_CONSTANTS[i];
  for (int i = 0; i < 12; i++) h.first_rc[i] = P2_PARTIAL_FIRST_RC[i];
  for (int i = 0; i < 22; i++) h.partial_rc[i] = P2_PARTIAL_RC[i];
  for (int i = 0; i < 242; i++) h.vs[i] = P2_PARTIAL_VS[i];
  for (int i = 0; i < 242; i++) h.w_hats[i] = P2_PARTIAL_W_HATS[i];
  for (int i = 0; i < 121; i++) h.init[i] = P2_PARTIAL_INIT_MATRIX[i];
  for (int k = 0; k < 8; k++)
    for (int i = 0; i < 12; i++) {
      u64 v;
      if (k < 3) v = P2_ROUND_CONSTANTS[12 * (k + 1) + i];
      else if (k == 3) v = P2_PARTIAL_FIRST_RC[i];
      else if (k < 7) v = P2_ROUND_CONSTANTS[12 * (26 + (k - 4) + 1) + i];
      else v = 0;
      h.post[12 * k + i] = v;
    }
  return cudaMemcpyToSymbol(C, &h, sizeof(h));
}

}  // namespace poseidon
