// Permutation argument: the Z polynomials and their partial products on the device.
// Reference (CPU): wires_permutation_partial_products_and_zs, plonky2/src/plonk/prover.rs:729-786;
// quotient_chunk_products / partial_products_and_z_gx, plonky2/src/util/partial_products.rs:13-37.
//
//   row i (x = w^i):  q_j = (wire_j + beta k_j x + gamma) / (wire_j + beta sigma_j + gamma),  j < num_routed
//   chunk products Q_k over `degree` consecutive j;  X_k = Q_0 ... Q_k
//   Z_0 = 1, Z_{i+1} = Z_i X_{K-1}(i);  partial product k of row i = Z_i X_k(i), k < K-1
// Output matrix (what the reference commits as zs_partial_products, prover.rs:112-117), column-major [nc*K][n]:
//   rows 0..nc-1 = Z_c, rows nc + c*(K-1) + k = partial product k of challenge c.
// Three kernels: per-row chunk products (one field inversion per row and challenge, Montgomery's trick across the
// chunks), CTA totals, and an in-CTA Hillis-Steele product scan that turns X into the final values in place.
#pragma once
#include "gl64.cuh"

namespace perm {
using gl::u64;
typedef uint32_t u32;
static constexpr int MAX_CH = 4;
static constexpr u32 SCAN_BLOCK = 1024;

struct Challenges {
  u64 beta[MAX_CH], gamma[MAX_CH];
};

// x^(p-2):  p - 2 = (2^32 - 2) * 2^32 + (2^32 - 1)
__device__ __forceinline__ u64 inverse(u64 x) {
  // a = x^(2^31 - 1) by the doubling chain 1, 2, 4, 8, 16 (+15) bits of ones
  auto sqn = [](u64 v, int n) {
    for (int i = 0; i < n; i++) v = gl::sqr(v);
    return v;
  };
  u64 x2 = gl::mul(sqn(x, 1), x);        // 2 ones
  u64 x4 = gl::mul(sqn(x2, 2), x2);      // 4
  u64 x8 = gl::mul(sqn(x4, 4), x4);      // 8
  u64 x16 = gl::mul(sqn(x8, 8), x8);     // 16
  u64 x24 = gl::mul(sqn(x16, 8), x8);    // 24
  u64 x28 = gl::mul(sqn(x24, 4), x4);    // 28
  u64 x30 = gl::mul(sqn(x28, 2), x2);    // 30
  u64 x31 = gl::mul(sqn(x30, 1), x);     // 31 ones = x^(2^31 - 1)
  u64 hi = gl::sqr(x31);                 // x^(2^32 - 2)
  u64 lo = gl::mul(hi, x);               // x^(2^32 - 1)
  return gl::mul(sqn(hi, 32), lo);
}

// thread = row.  out as described above; X_k is left in the partial-product slots and X_{K-1} in the Z slot.
template <int MAXK>
__global__ void __launch_bounds__(128) chunk_products_kernel(const u64* __restrict__ wires, const u64* __restrict__ sigmas,
                                                             u64 n, u32 n_log, u32 num_routed, u32 degree, u32 nc, Challenges ch,
                                                             const u64* __restrict__ k_is, u64 w, u64* __restrict__ out,
                                                             u32* __restrict__ zero_denominator) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 K = (num_routed + degree - 1) / degree;
  const u64 x = gl::pow(w, i);
  for (u32 c = 0; c < nc; c++) {
    const u64 beta = ch.beta[c], gamma = ch.gamma[c];
    const u64 bx = gl::mul(beta, x);
    u64 N[MAXK], D[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; k++) {
      N[k] = 1;
      D[k] = 1;
      if ((u32)k < K) {
        const u32 j1 = min(num_routed, (u32)(k + 1) * degree);
        for (u32 j = (u32)k * degree; j < j1; j++) {
          const u64 wv = __ldg(wires + (u64)j * n + i);
          N[k] = gl::mul(N[k], gl::add(gl::mul_add(bx, __ldg(k_is + j), wv), gamma));
          D[k] = gl::mul(D[k], gl::add(gl::mul_add(beta, __ldg(sigmas + (u64)j * n + i), wv), gamma));
        }
      }
    }
    // 1 / D_k for all chunks with one inversion
    u64 pd[MAXK];
    u64 run = 1;
#pragma unroll
    for (int k = 0; k < MAXK; k++) {
      pd[k] = run;  // product of D_0..D_{k-1}
      run = gl::mul(run, D[k]);
    }
    // the reference's batch_multiplicative_inverse panics on a zero denominator ("Tried to invert zero", types.rs:130): flagged
    // here, turned into an error by the host entry point
    if (gl::canon(run) == 0) atomicOr(zero_denominator, 1u);
    u64 inv = inverse(run);
#pragma unroll
    for (int k = MAXK - 1; k >= 0; k--) {
      u64 dinv = gl::mul(inv, pd[k]);
      inv = gl::mul(inv, D[k]);
      N[k] = gl::mul(N[k], dinv);  // Q_k
    }
    u64 acc = 1;
#pragma unroll
    for (int k = 0; k < MAXK; k++) {
      if ((u32)k < K) {
        acc = gl::mul(acc, N[k]);  // X_k
        u64* dst = (u32)k + 1 < K ? out + ((u64)nc + (u64)c * (K - 1) + k) * n : out + (u64)c * n;
        dst[i] = gl::canon(acc);
      }
    }
  }
}

// totals[c][b] = product of the Z-slot values (X_{K-1}) of CTA b's rows
__global__ void __launch_bounds__(SCAN_BLOCK) block_totals_kernel(const u64* __restrict__ out, u64 n, u32 nb, u64* __restrict__ totals) {
  __shared__ u64 sh[SCAN_BLOCK];
  const u32 t = threadIdx.x, c = blockIdx.y;
  const u64 i = (u64)blockIdx.x * SCAN_BLOCK + t;
  sh[t] = i < n ? out[(u64)c * n + i] : 1;
  __syncthreads();
  for (u32 d = SCAN_BLOCK / 2; d > 0; d >>= 1) {
    if (t < d) sh[t] = gl::mul(sh[t], sh[t + d]);
    __syncthreads();
  }
  if (t == 0) totals[(u64)c * nb + blockIdx.x] = sh[0];
}

// totals[c][b] <- product of totals[c][0..b)   (exclusive; one CTA per challenge, thread q owns G consecutive CTAs)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_totals_kernel(u64* __restrict__ totals, u32 nb) {
  __shared__ u64 sh[SCAN_BLOCK];
  const u32 q = threadIdx.x;
  u64* T = totals + (u64)blockIdx.x * nb;
  const u32 G = (nb + SCAN_BLOCK - 1) / SCAN_BLOCK;
  const u32 b0 = min(nb, q * G), b1 = min(nb, b0 + G);
  u64 acc = 1;
  for (u32 b = b0; b < b1; b++) acc = gl::mul(acc, T[b]);
  sh[q] = acc;
  __syncthreads();
  for (u32 d = 1; d < SCAN_BLOCK; d <<= 1) {  // inclusive prefix products
    u64 v = sh[q];
    if (q >= d) v = gl::mul(v, sh[q - d]);
    __syncthreads();
    sh[q] = v;
    __syncthreads();
  }
  u64 cur = q ? sh[q - 1] : 1;
  for (u32 b = b0; b < b1; b++) {
    u64 h = T[b];
    T[b] = cur;
    cur = gl::mul(cur, h);
  }
}

// Z_i = carry[CTA] * prod_{i' < i in the CTA} X_{K-1}(i');  Z slot <- Z_i;  partial-product slots *= Z_i
__global__ void __launch_bounds__(SCAN_BLOCK) apply_kernel(u64* __restrict__ out, u64 n, u32 nb, u32 nc, u32 K,
                                                           const u64* __restrict__ totals) {
  __shared__ u64 sh[SCAN_BLOCK];
  const u32 t = threadIdx.x, c = blockIdx.y;
  const u64 i = (u64)blockIdx.x * SCAN_BLOCK + t;
  sh[t] = i < n ? out[(u64)c * n + i] : 1;
  __syncthreads();
  for (u32 d = 1; d < SCAN_BLOCK; d <<= 1) {
    u64 v = sh[t];
    if (t >= d) v = gl::mul(v, sh[t - d]);
    __syncthreads();
    sh[t] = v;
    __syncthreads();
  }
  if (i >= n) return;
  u64 z = totals[(u64)c * nb + blockIdx.x];
  if (t) z = gl::mul(z, sh[t - 1]);
  out[(u64)c * n + i] = gl::canon(z);
  for (u32 k = 0; k + 1 < K; k++) {
    u64* p = out + ((u64)nc + (u64)c * (K - 1) + k) * n + i;
    *p = gl::canon(gl::mul(*p, z));
  }
}

}  // namespace perm
