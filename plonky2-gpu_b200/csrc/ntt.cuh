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
// checks the CTA-wide OR of m.any() before storing and, if set (~2^-32 per butterfly), reloads its tile and redoes it
// with the exact form, so results are bit-exact for every input.
template <class M>
__device__ __forceinline__ void butterfly(u64& u, u64& v, u64 z, M& m) {
  u64 t = gl::mul(z, v, m);
  u64 a = gl::add(u, t, m);
  v = gl::sub(u, t, m);
  u = a;
}

// to_canonical_u64 on the carry chain: x >= p  <=>  x + eps wraps, and then x - p == x + eps (mod 2^64)
__device__ __forceinline__ u64 canon_fast(u64 x) {
  u64 t;
  u32 c;
  asm("{ add.cc.u64 %0, %2, 0xffffffff; addc.u32 %1, 0, 0; }" : "=l"(t), "=r"(c) : "l"(x));
  return c ? t : x;
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
// twiddle accessors: tw(level, block) for local level / local block of the tile
struct StagedTwiddles {  // W[(1 << level) - 1 + block] in shared memory (one evaluation-tree node per CTA)
  const u64* W;
  __device__ __forceinline__ u64 operator()(int level, int block) const { return W[(1 << level) - 1 + block]; }
};
struct GlobalTwiddles {  // U[(q << level) + block] through L1: every lane of the tile is a different tree node q
  const u64* U;
  u64 q;
  __device__ __forceinline__ u64 operator()(int level, int block) const { return __ldg(U + ((q << level) + block)); }
};

template <int R, class M, class TW>
__device__ __forceinline__ void radix_group(u64* __restrict__ x, const TW& tw, int L, int a, int T,
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
        u64 z = tw(a + u, (lh << u) + (m >> (R - u)));
        butterfly(v[m], v[m + half], z, mode);
      }
    }
#pragma unroll
    for (int m = 0; m < (1 << R); m++) x[(base_l + (m << ll_bits)) * TP + t] = v[m];
  }
}

// all L levels of the tile; ends with a __syncthreads
template <class M, class TW>
__device__ __forceinline__ void run_levels(u64* x, const TW& W, int L, int T, int TP, int tid, int nthreads, M& m) {
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
  // thread -> (row l0 + k * rows_per_iter, position t): the pointer advances by a fixed stride per iteration, so the
  // copy loops are a load/store, a pointer add and a compare per element (nt is a multiple of T)
  const int t = tid % T, l0 = tid / T, rows_per_iter = nt / T;
  const u64 step = (u64)rows_per_iter * stride;
  {
    const u64* sp = src + (u64)l0 * stride + t;
    for (int i = tid; i < (T << L); i += nt, sp += step) x[i] = *sp;
  }
  __syncthreads();
  gl::Optimistic fast;
  run_levels(x, StagedTwiddles{W}, L, T, T, tid, nt, fast);
  if (__syncthreads_or(fast.any() | (p.force_redo != 0))) {  // the source tile is still intact (nothing stored yet): redo it exactly
    const u64* sp = src + (u64)l0 * stride + t;
    for (int i = tid; i < (T << L); i += nt, sp += step) x[i] = *sp;
    __syncthreads();
    gl::Exact exact;
    run_levels(x, StagedTwiddles{W}, L, T, T, tid, nt, exact);
  }
  {
    u64* dp = dst + (u64)l0 * stride + t;
    for (int i = tid; i < (T << L); i += nt, dp += step) *dp = x[i];
  }
}

// ---- K2: final pass (s0 + L == k), chunks of 2^L contiguous positions ---------------------------------
// MODE_COLMAJOR : dst column-major, same positions (bit-reversed order), canonical values.
// MODE_ROWS     : dst row-major rows: dst[(row0 + q*2^L + l) * row_stride + col0 + c]  (LDE leaves)
// CTA = one chunk q x C columns; smem x[2^L][C+1] + W[2^L].  grid.x = 2^s0 chunks * ceil(ncols / C) column groups, the
// column group varying fastest: the CTAs that write the C-column pieces of the SAME leaf rows are resident together, so
// the 64-byte pieces (not sector-aligned when the row stride is odd) merge into whole lines in L2 before they reach HBM
// (chunk-fastest order: 2.13 GB read + 1.54 GB written per 1.13 GB coset block, profiles/r02b_ntt.md; LDE of 2^20 x 135:
// 26.6 -> 22.9 ms).
enum { MODE_COLMAJOR = 0, MODE_ROWS = 1 };
template <int C, int MODE>
__global__ void __launch_bounds__(512) ntt_final_pass_kernel(PassArgs p, LevelScale sc, u64 row0, u64 row_stride,
                                                             u64 col0) {
  extern __shared__ __align__(16) u64 smem[];
  const int L = p.L;
  constexpr int TP = C + 1;
  u64* x = smem;
  u64* W = smem + ((size_t)TP << L);
  const u32 ngroups = (p.ncols + C - 1) / C;
  const u64 q = blockIdx.x / ngroups;
  const u32 c0 = (blockIdx.x % ngroups) * C;
  const int nc = min((u32)C, p.ncols - c0);
  const int tid = threadIdx.x, nt = blockDim.x;
  const u64 base = q << L;

  stage_twiddles(W, p.U, sc, p.scaled, p.s0, L, (p.block_base << p.s0) + q, tid, nt);
  auto load_tile = [&]() {
    // consecutive threads read consecutive positions of a column (coalesced); the C column loads of a position are
    // issued back to back (predicated, independent) before anything is stored to shared memory
    const u64* sp = p.src + (u64)c0 * p.src_cs + base;
    for (int l = tid; l < (1 << L); l += nt) {
      u64 v[C];
#pragma unroll
      for (int c = 0; c < C; c++) v[c] = c < nc ? sp[(u64)c * p.src_cs + l] : 0;
#pragma unroll
      for (int c = 0; c < C; c++) x[l * TP + c] = v[c];
    }
  };
  load_tile();
  __syncthreads();
  gl::Optimistic fast;
  run_levels(x, StagedTwiddles{W}, L, C, TP, tid, nt, fast);
  if (__syncthreads_or(fast.any() | (p.force_redo != 0))) {
    load_tile();
    __syncthreads();
    gl::Exact exact;
    run_levels(x, StagedTwiddles{W}, L, C, TP, tid, nt, exact);
  }
  if (MODE == MODE_COLMAJOR) {
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (c < nc) {
        u64* dp = p.dst + (u64)(c0 + c) * p.dst_cs + base;
        for (int l = tid; l < (1 << L); l += nt) dp[l] = canon_fast(x[l * TP + c]);
      }
    }
  } else {
    // leaf rows: thread -> (row l0 + k * rows_per_iter, column c): C consecutive threads write one row's C columns
    const int c = tid % C, l0 = tid / C, rows_per_iter = nt / C;
    if (c < nc) {
      u64* dp = p.dst + (row0 + base + l0) * row_stride + col0 + c0 + c;
      const u64 step = (u64)rows_per_iter * row_stride;
      for (int l = l0; l < (1 << L); l += rows_per_iter, dp += step) *dp = canon_fast(x[l * TP + c]);
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

  // load: chunk j is contiguous in global memory (coalesced along l); eight independent loads per thread are in flight
  // before the first one is stored (the rolled form waited on every load: long-scoreboard 4.5 per issue, profiles/r02b_ntt.md)
  for (int i0 = tid; i0 < (J << L); i0 += 8 * nt) {
    u64 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = i0 + u * nt, l = i & ((1 << L) - 1), j = i >> L;
      v[u] = 0;
      if (i < (J << L) && j < nj) {
        u64 q = s0 ? (__brevll(qp0 + j) >> (64 - s0)) : 0;
        v[u] = src[(q << L) + l];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = i0 + u * nt, l = i & ((1 << L) - 1), j = i >> L;
      if (i < (J << L) && j < nj) x[l * J + j] = v[u];
    }
  }
  __syncthreads();
  // levels: the shared radix-8/4/2 machinery with per-lane twiddles U[(q_j << a) + lb] (lane j = tid % J is fixed
  // per thread because J divides the block size); exact reductions (the source chunks are re-read nowhere else)
  {
    const int j = tid % J;
    GlobalTwiddles tw{p.U, s0 ? (__brevll(qp0 + min(j, nj - 1)) >> (64 - s0)) : 0};
    gl::Exact exact;
    run_levels(x, tw, L, J, J, tid, nt, exact);
  }
  // store: i = rev_L(l) * 2^s0 + q'  -> for fixed l, J consecutive outputs
  for (int i = tid; i < (J << L); i += nt) {
    int j = i % J, l = i / J;
    if (j < nj) {
      u64 hi = L ? (u64)(__brev((u32)l) >> (32 - L)) : 0;
      dst[(hi << s0) + qp0 + j] = canon_fast(gl::mul(x[l * J + j], n_inv));
    }
  }
}

}  // namespace ntt
