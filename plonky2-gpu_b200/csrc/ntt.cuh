// ntt.cuh -- batched Goldilocks NTTs for the commit path, as shared-memory radix-2^k passes.
//
// What is computed (bit-exact with the reference CPU path):
//   * inverse NTT of each column            PolynomialValues::ifft -> ifft_with_options, field/src/fft.rs:73-103
//   * coset low-degree extension            PolynomialCoeffs::lde + coset_fft_with_options(shift = 7,
//                                            zero_factor = rate_bits), field/src/polynomial/mod.rs:205-224, 286-299
//     delivered directly in the reference's leaf order (transpose + reverse_index_bits_in_place,
//     fri/oracle.rs:942-952) with no separate bit-reversal or transpose pass.
//
// Formulation (why there is no bit-reversal kernel).  Evaluate f (n = 2^k coefficients, natural order) on
// the coset  S * <w_n>  with the "evaluation tree": a node at level s, block B holds a polynomial of
// n / 2^s coefficients to be evaluated on the coset  S * w^{rev_s(B)} * <w^{2^s}>;  one butterfly level
//      (lo, hi) -> (lo + z*hi, lo - z*hi),   z = z(s, B) = S^{n/2^{s+1}} * U[B],
//      U[B] = prod_{m in bits(B)} root(m + 2)  ( = w_{2^{s+1}}^{rev_s(B)} for every s > log2 B ),
// splits it into its two children, and after k levels position j holds f(S * w^{rev_k(j)}): natural-order
// input, bit-reversed output, twiddles constant inside a block and shared by every column.  The table U is
// the same for every n (prefix property of root(k)^2 = root(k-1), field/src/types.rs:268-272).
//
//   LDE (rate 2^r, N = n * 2^r, S = g = 7): the zero-padded size-N tree starts with r trivial levels that
//   copy the coefficients into 2^r blocks; block b is the size-n tree with S_b = g * w_N^{rev_r(b)}, i.e.
//   z(s, B) = g^{N / 2^{r+s+1}} * U[(b << s) + B] -- levels r..r+k-1 of ONE universal table.  Leaf row
//   L = b*n + j of the reference is exactly position j of block b.
//
//   iNTT: coefficient c_i = n^{-1} * sum_m v_m w^{-im}: the same tree with U_inv (inverse roots), S = 1,
//   whose output position j holds n * c_{rev_k(j)}; the final pass stores transposed tiles so that the
//   scatter to natural order stays 128-byte coalesced, and multiplies by n^{-1} on the way out.
//
// Pass structure: a size-2^k transform is cut into passes of L <= MAX_L levels.  A pass loads a tile
// [2^L][T] into shared memory (T = adjacent independent positions, so global accesses are T*8-byte
// segments), runs its levels as radix-8/4/2 register butterflies with one __syncthreads per radix group,
// and stores.  Twiddles of the tile (2^L - 1 values, sigma pre-multiplied) are staged in shared memory.
#pragma once
#include "gl64.cuh"

namespace ntt {

using gl::u32;
using gl::u64;

static constexpr int MAX_LEVELS = 34;

struct LevelScale {
  u64 sigma[MAX_LEVELS];  // sigma[s] multiplies U at level s of the sub-problem (all ones for a plain NTT)
};

// (u, v) <- (u + z v, u - z v).  M: exact or optimistic reduction (gl64.cuh); a kernel that runs the optimistic form
// checks the CTA-wide OR of m.rare before storing and, if set (~2^-32 per butterfly), reloads its tile and redoes it
// with the exact form, so results are bit-exact for every input.
template <class M>
__device__ __forceinline__ void butterfly(u64& u, u64& v, u64 z, M& m) {
  u64 t = gl::mul(z, v, m);
  u64 a = gl::add(u, t);
  v = gl::sub(u, t);
  u = a;
}

// U[b] = prod_{m in bits(b)} root(m + 2);  roots[j] = primitive_root_of_unity(j) (or its inverse).
__global__ void build_twiddles_kernel(u64* __restrict__ U, u64 count, const u64* __restrict__ roots) {
  u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= count) return;
  u64 acc = 1;
  u64 x = b;
  for (int m = 0; x; m++, x >>= 1)
    if (x & 1) acc = gl::mul(acc, roots[m + 2]);
  U[b] = gl::canon(acc);
}

// ---- in-shared-memory levels --------------------------------------------------------------------------
// x: [2^L][TP] (t fastest, TP >= T), W: twiddles of the tile, W[(1 << a) - 1 + lb] for local level a, local
// block lb.  Runs local levels [a, a + R) with R in {1,2,3} in registers.  All threads must call.
template <int R, class M>
__device__ __forceinline__ void radix_group(u64* __restrict__ x, const u64* __restrict__ W, int L, int a, int T,
                                            int TP, int tid, int nthreads, M& mode) {
  const int ll_bits = L - a - R;
  const int items = (1 << (L - R)) * T;
  for (int it = tid; it < items; it += nthreads) {
    int t = it % T;
    int w = it / T;
    int ll = w & ((1 << ll_bits) - 1);
    int lh = w >> ll_bits;
    int base_l = (lh << (L - a)) + ll;
    u64 v[1 << R];
#pragma unroll
    for (int m = 0; m < (1 << R); m++) v[m] = x[(base_l + (m << ll_bits)) * TP + t];
#pragma unroll
    for (int u = 0; u < R; u++) {
      const int half = 1 << (R - 1 - u);
#pragma unroll
      for (int m = 0; m < (1 << R); m++) {
        if (m & half) continue;
        u64 z = W[(1 << (a + u)) - 1 + (lh << u) + (m >> (R - u))];
        butterfly(v[m], v[m + half], z, mode);
      }
    }
#pragma unroll
    for (int m = 0; m < (1 << R); m++) x[(base_l + (m << ll_bits)) * TP + t] = v[m];
  }
}

// all L levels of the tile; ends with a __syncthreads
template <class M>
__device__ __forceinline__ void run_levels(u64* x, const u64* W, int L, int T, int TP, int tid, int nthreads, M& m) {
  int a = 0;
  while (a < L) {
    int rem = L - a;
    if (rem >= 3 && rem != 4) {
      radix_group<3>(x, W, L, a, T, TP, tid, nthreads, m);
      a += 3;
    } else if (rem >= 2) {
      radix_group<2>(x, W, L, a, T, TP, tid, nthreads, m);
      a += 2;
    } else {
      radix_group<1>(x, W, L, a, T, TP, tid, nthreads, m);
      a += 1;
    }
    __syncthreads();
  }
}

// stage the tile's twiddles: W[(1<<a)-1+lb] = sigma[s0+a] * U[(G0 << a) + lb]
__device__ __forceinline__ void stage_twiddles(u64* W, const u64* __restrict__ U, const LevelScale& sc, bool scaled,
                                               int s0, int L, u64 G0, int tid, int nthreads) {
  for (int i = tid; i < (1 << L) - 1; i += nthreads) {
    int a = 31 - __clz(i + 1);
    int lb = i + 1 - (1 << a);
    u64 z = __ldg(U + ((G0 << a) + lb));
    if (scaled) z = gl::mul(z, sc.sigma[s0 + a]);
    W[i] = z;
  }
}

struct PassArgs {
  const u64* src;   // column-major: element i of column c at src[c * src_cs + i]
  u64* dst;
  u64 src_cs, dst_cs;
  u32 k;            // log2 of the transform size
  u32 s0, L;        // this pass runs levels [s0, s0 + L)
  u64 block_base;   // LDE coset block b: block index at level s is (block_base << s) + B
  const u64* U;
  u32 scaled;       // use sc.sigma
  u32 ncols;
  u32 force_redo;   // test hook: take the exact-redo path as if an optimistic reduction had reported a rare case
};

// ---- K1: strided pass (s0 + L < k).  grid.x = 2^s0 * (stride / T) tiles, grid.y = columns -------------
// dynamic smem: (2^L * T + 2^L) * 8 bytes
template <int T>
__global__ void __launch_bounds__(512) ntt_strided_pass_kernel(PassArgs p, LevelScale sc) {
  extern __shared__ __align__(16) u64 smem[];
  const int L = p.L;
  u64* x = smem;
  u64* W = smem + ((size_t)T << L);
  const u64 stride = ((u64)1 << p.k) >> (p.s0 + L);
  const u64 tiles_per_block = stride / T;
  const u64 Bhi = blockIdx.x / tiles_per_block;
  const u64 j0 = (blockIdx.x % tiles_per_block) * T;
  const u64 col = blockIdx.y;
  const u64 base = Bhi * (((u64)1 << p.k) >> p.s0) + j0;
  const u64* src = p.src + col * p.src_cs + base;
  u64* dst = p.dst + col * p.dst_cs + base;
  const int tid = threadIdx.x, nt = blockDim.x;

  stage_twiddles(W, p.U, sc, p.scaled, p.s0, L, (p.block_base << p.s0) + Bhi, tid, nt);
  for (int i = tid; i < (T << L); i += nt) {
    int t = i % T, l = i / T;
    x[i] = src[(u64)l * stride + t];
  }
  __syncthreads();
  gl::Optimistic fast;
  run_levels(x, W, L, T, T, tid, nt, fast);
  if (__syncthreads_or(fast.rare | (p.force_redo != 0))) {  // the source tile is still intact (nothing stored yet): redo it exactly
    for (int i = tid; i < (T << L); i += nt) {
      int t = i % T, l = i / T;
      x[i] = src[(u64)l * stride + t];
    }
    __syncthreads();
    gl::Exact exact;
    run_levels(x, W, L, T, T, tid, nt, exact);
  }
  for (int i = tid; i < (T << L); i += nt) {
    int t = i % T, l = i / T;
    dst[(u64)l * stride + t] = x[i];
  }
}

// ---- K2: final pass (s0 + L == k), chunks of 2^L contiguous positions ---------------------------------
// MODE_COLMAJOR : dst column-major, same positions (bit-reversed order), canonical values.
// MODE_ROWS     : dst row-major rows: dst[(row0 + q*2^L + l) * row_stride + col0 + c]  (LDE leaves)
// CTA = one chunk q x C columns; smem x[2^L][C+1] + W[2^L].  grid.x = 2^s0 chunks, grid.y = ceil(ncols / C).
enum { MODE_COLMAJOR = 0, MODE_ROWS = 1 };
template <int C, int MODE>
__global__ void __launch_bounds__(512) ntt_final_pass_kernel(PassArgs p, LevelScale sc, u64 row0, u64 row_stride,
                                                             u64 col0) {
  extern __shared__ __align__(16) u64 smem[];
  const int L = p.L;
  constexpr int TP = C + 1;
  u64* x = smem;
  u64* W = smem + ((size_t)TP << L);
  const u64 q = blockIdx.x;
  const u32 c0 = blockIdx.y * C;
  const int nc = min((u32)C, p.ncols - c0);
  const int tid = threadIdx.x, nt = blockDim.x;
  const u64 base = q << L;

  stage_twiddles(W, p.U, sc, p.scaled, p.s0, L, (p.block_base << p.s0) + q, tid, nt);
  for (int i = tid; i < (C << L); i += nt) {
    int l = i & ((1 << L) - 1), c = i >> L;
    x[l * TP + c] = c < nc ? p.src[(u64)(c0 + c) * p.src_cs + base + l] : 0;
  }
  __syncthreads();
  gl::Optimistic fast;
  run_levels(x, W, L, C, TP, tid, nt, fast);
  if (__syncthreads_or(fast.rare | (p.force_redo != 0))) {
    for (int i = tid; i < (C << L); i += nt) {
      int l = i & ((1 << L) - 1), c = i >> L;
      x[l * TP + c] = c < nc ? p.src[(u64)(c0 + c) * p.src_cs + base + l] : 0;
    }
    __syncthreads();
    gl::Exact exact;
    run_levels(x, W, L, C, TP, tid, nt, exact);
  }
  if (MODE == MODE_COLMAJOR) {
    for (int i = tid; i < (C << L); i += nt) {
      int l = i & ((1 << L) - 1), c = i >> L;
      if (c < nc) p.dst[(u64)(c0 + c) * p.dst_cs + base + l] = gl::canon(x[l * TP + c]);
    }
  } else {
    for (int i = tid; i < (C << L); i += nt) {
      int c = i % C, l = i / C;
      if (c < nc) p.dst[(row0 + base + l) * row_stride + col0 + c0 + c] = gl::canon(x[l * TP + c]);
    }
  }
}

// ---- K2b: final pass of the inverse NTT: natural-order, scaled output --------------------------------
// CTA = J chunks whose bit-reversed chunk indices are consecutive (q = rev_{s0}(q'), q' in [J*blockIdx.x, +J))
// of one column; after the levels, position l of chunk q' holds n * c_i with i = rev_L(l) * 2^s0 + q'.
// Twiddles differ per chunk, so they are read through L1 from the global table instead of staged.
template <int J>
__global__ void __launch_bounds__(512) intt_final_pass_kernel(PassArgs p, u64 n_inv) {
  extern __shared__ __align__(16) u64 smem[];
  const int L = p.L;
  u64* x = smem;  // [2^L][J]  (j fastest)
  const int tid = threadIdx.x, nt = blockDim.x;
  const u32 s0 = p.s0;
  const u64 qp0 = (u64)blockIdx.x * J;
  const u64 col = blockIdx.y;
  const u64* src = p.src + col * p.src_cs;
  u64* dst = p.dst + col * p.dst_cs;
  const int nj = (int)min((u64)J, ((u64)1 << s0) - qp0);  // s0 < log2 J: fewer chunks than J

  // load: chunk j is contiguous in global memory (coalesced along l)
  for (int i = tid; i < (J << L); i += nt) {
    int l = i & ((1 << L) - 1), j = i >> L;
    if (j < nj) {
      u64 q = s0 ? (__brevll(qp0 + j) >> (64 - s0)) : 0;
      x[l * J + j] = src[(q << L) + l];
    }
  }
  __syncthreads();
  // levels: radix-2 groups with per-lane twiddles U[(q_j << a) + lb]
  for (int a = 0; a < L; a++) {
    const int ll_bits = L - a - 1;
    const int items = (1 << (L - 1)) * J;
    for (int it = tid; it < items; it += nt) {
      int j = it % J, w = it / J;
      if (j >= nj) continue;
      int ll = w & ((1 << ll_bits) - 1), lh = w >> ll_bits;
      int l0 = (lh << (L - a)) + ll, l1 = l0 + (1 << ll_bits);
      u64 q = s0 ? (__brevll(qp0 + j) >> (64 - s0)) : 0;
      u64 z = __ldg(p.U + ((q << a) + lh));
      u64 u = x[l0 * J + j], v = x[l1 * J + j];
      gl::Exact exact;
      butterfly(u, v, z, exact);
      x[l0 * J + j] = u;
      x[l1 * J + j] = v;
    }
    __syncthreads();
  }
  // store: i = rev_L(l) * 2^s0 + q'  -> for fixed l, J consecutive outputs
  for (int i = tid; i < (J << L); i += nt) {
    int j = i % J, l = i / J;
    if (j < nj) {
      u64 hi = L ? (u64)(__brev((u32)l) >> (32 - L)) : 0;
      dst[(hi << s0) + qp0 + j] = gl::canon(gl::mul(x[l * J + j], n_inv));
    }
  }
}

}  // namespace ntt
