// poseidon.cuh -- Poseidon permutation over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds) and the
// overwrite-mode sponge / 2-to-1 compression built on it, for sm_100a.
//
// What is computed (bit-exact with the reference CPU path):
//   permute()        Poseidon::poseidon            plonky2/src/hash/poseidon.rs:590-606, evaluated in the algebraically
//                                                  identical "naive" round structure poseidon_naive (poseidon.rs:619-640):
//                                                  30 x [constant_layer, S-box (all lanes | lane 0), mds_layer]
//   S-box            sbox_monomial                 poseidon.rs:525-548
//   MDS layer        mds_row_shf / mds_layer       poseidon.rs:172-260
//   fast partial rounds (kept for the Poseidon GATE evaluator, whose wires are defined in that basis, quotient.cuh)
//                                                  poseidon.rs:310-365, 398-427, 574-588
//   hash_or_noop / hash_no_pad / two_to_one        plonk/config.rs:56-67, hash/hashing.rs:81-104, :65-72
//
// How (B200).  Measured on the box (profiles/r02_pipe_model.md): IMAD.WIDE.U32 costs ~4.3 issue cycles per warp on the
// FMA-heavy pipe and cannot overlap FP64 instructions; IADD3/LOP3 (ALU), 32-bit IMAD and DADD/DFMA (FP64, 64 lanes/clk/SM
// on B200) cost ~2 cycles each on three pipes that do overlap.  The reference's fast partial rounds trade one small-
// constant MDS layer for 22 dense 64x64-bit multiplications, which is the right trade on a CPU and the wrong one here:
// a dense multiplication is 6 IMAD.WIDE, while the MDS layer needs NO multiplier at all -- its circulant becomes, by the
// Chinese remainder theorem over t^12 - 1, 78 additions / multiply-adds by +-2^k (mds_fft.cuh), run exactly on the FP64
// pipe over the 32-bit halves of the state words.  So every round is evaluated in the naive form: S-boxes as 64-bit
// modular multiplications (gl64.cuh, optimistic reduction with an exact redo), the linear layer in doubles, the next
// round's constants folded into the double -> integer conversion.  Per permutation: 16.1 k thread-instructions
// (2.9 k IMAD.WIDE, 5.5 k FP64) against 21.8 k (8.7 k IMAD.WIDE) for the fast form of round 1.
#pragma once
#include <string.h>
#include "gl64.cuh"
#include "mds_fft.cuh"
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
  double rc_d[31 * 12 * 2];   // 2^52 + low half, 2^52 + high half of ALL_ROUND_CONSTANTS row r (row 30 = zeros): the constant
                              // layer of round r rides on the double -> integer conversion after the MDS of round r - 1
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

// (2^52 + L, 2^52 + H) with integers 0 <= L, H < 2^51  ->  a u64 representative of L + 2^32 * H  (mod p).
// The integers sit in the mantissas: low word = bits 0..31, high word = 0x43300000 + bits 32..51.
__device__ __forceinline__ u64 from_biased_halves(double al, double ah) {
  u32 w0 = (u32)__double2loint(al), l1 = (u32)__double2hiint(al), h0 = (u32)__double2loint(ah), h1 = (u32)__double2hiint(ah), w1, w2;
  asm("{ .reg .u32 t, u; sub.u32 t, %2, 0x43300000; add.cc.u32 %0, t, %3; sub.u32 u, %4, 0x43300000; addc.u32 %1, u, 0; }"
      : "=r"(w1), "=r"(w2) : "r"(l1), "r"(h0), "r"(h1));
  return gl::reduce96(gl::pack(w0, w1), w2);
}

// Optimistic form without any multiply: with L = l0 + 2^32 l1, H = h0 + 2^32 h1 (l1, h1 < 2^19) and 2^64 == 2^32 - 1,
//   L + 2^32 H  ==  (l0 - h1) + 2^32 (h0 + l1 + h1)   (mod p);
// the middle sum overflows 32 bits with probability ~2^-22 -- that case is flagged in m.rare and the caller redoes its
// work with the exact functions (the borrow of l0 - h1 is propagated exactly; it can only underflow the high word
// together with the flagged overflow).  6 ALU instructions instead of 2 IMAD.WIDE + 4.
template <class M>
__device__ __forceinline__ u64 from_biased_halves(double al, double ah, M& m) {
  if (M::optimistic) {
    u32 l0 = (u32)__double2loint(al), hl = (u32)__double2hiint(al), h0 = (u32)__double2loint(ah), hh = (u32)__double2hiint(ah);
    u32 u = hl + hh - 0x86600000u;   // l1 + h1
    u32 h1 = hh - 0x43300000u;
    u32 mid = h0 + u;
    m.rare |= mid < u;
    u32 lo, hi;
    asm("{ sub.cc.u32 %0, %2, %3; subc.u32 %1, %4, 0; }" : "=r"(lo), "=r"(hi) : "r"(l0), "r"(h1), "r"(mid));
    return gl::pack(lo, hi);
  }
  return from_biased_halves(al, ah);
}

// out[r] = sum_i circ[i] * s[(i+r)%12] + diag[r]*s[r] + addc[r]   (mds_row_shf + mds_layer, then the next
// constant_layer folded in).  addc must be canonical round constants (< p).
__device__ __forceinline__ void mds_layer(u64 (&s)[12], const u64* __restrict__ addc) {
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
}

// sbox_layer (poseidon.rs:534-548).  Two forms: the hash kernels inline all 12 S-boxes (no register moves; their round
// loop is small since the MDS layer shrank to ~350 instructions); the gate evaluator (quotient.cuh), whose kernel is
// instruction-cache bound, uses the rolled 3 x 4 form with a register rotation.
template <class M>
__device__ __forceinline__ void sbox_layer(u64 (&s)[12], M& m) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = sbox(s[i], m);
}
template <class M>
__device__ __forceinline__ void sbox_layer_rolled(u64 (&s)[12], M& m) {
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

// mds_partial_layer_init (poseidon.rs:310-337): s[0] unchanged; s[c] = sum_r s[r] * init[r-1][c-1].
// Rolled over the output column c (keeps ~24 KB of straight-line code out of the instruction cache); the results
// are pushed through a shift register so no register array is indexed dynamically.
template <class M>
__device__ __forceinline__ void partial_layer_init(u64 (&s)[12], M& m) {
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

// ---- the permutation ---------------------------------------------------------------------------------------------------
// One MDS layer in the naive round structure + the NEXT round's constant layer: bias[2r], bias[2r+1] = 2^52 + the halves of
// that constant, so "add the constant" and "move the integer into the mantissa" are the same DADD.
template <class M>
__device__ __forceinline__ void mds_naive(u64 (&s)[12], const double* __restrict__ bias, M& m) {
  double lo[12], hi[12], yl[12], yh[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u32 l, h;
    gl::split(s[i], l, h);
    lo[i] = __uint2double_rn(l);   // I2F.F64.U32: exact, and on the otherwise idle XU pipe
    hi[i] = __uint2double_rn(h);
  }
  mdsfft::mds12<double>(lo, yl);
  mdsfft::mds12<double>(hi, yh);
#pragma unroll
  for (int r = 0; r < 12; r++) s[r] = from_biased_halves(yl[r] + bias[2 * r], yh[r] + bias[2 * r + 1], m);
}

// Input: any u64 representatives; output: u64 representatives (NOT canonicalised -- callers canonicalise what they store).
template <class M>
__device__ __forceinline__ void permute(u64 (&s)[12], M& m) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_canonical(s[i], C.rc[i]);
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
#ifndef P2B_LAB_NOSBOX   // (timing experiments only: tools/poseidon_lab.cu)
    if (r < 4 || r >= 26) sbox_layer(s, m);
    else s[0] = sbox(s[0], m);
#endif
#ifndef P2B_LAB_NOMDS
    mds_naive(s, C.rc_d + 24 * (r + 1), m);
#endif
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
  auto biased = [](u32 half) {
    u64 bits = 0x4330000000000000ull | (u64)half;  // 2^52 + half, exactly
    double d;
    memcpy(&d, &bits, 8);
    return d;
  };
  for (int i = 0; i < 31 * 12; i++) {
    u64 v = i < 360 ? P2_ROUND_CONSTANTS[i] : 0;
    h.rc_d[2 * i] = biased((u32)v);
    h.rc_d[2 * i + 1] = biased((u32)(v >> 32));
  }
  return cudaMemcpyToSymbol(C, &h, sizeof(h));
}

}  // namespace poseidon
