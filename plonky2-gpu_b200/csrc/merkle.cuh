// merkle.cuh -- Poseidon Merkle tree over leaf rows: leaf hashing, digest layers, cap.
//
// Reference semantics: MerkleTree::new / fill_digests_buf / fill_subtree
// (plonky2/src/hash/merkle_tree.rs:283-319, 210-244, 78-105); leaf digest = hash_or_noop
// (plonk/config.rs:56-67) over the overwrite-mode sponge hash_n_to_m_no_pad (hash/hashing.rs:81-104);
// inner node = two_to_one (hashing.rs:65-72).
//
// Digest layout (merkle_tree.rs:46-54, 424-435): `digests` = 2^cap_height sub-trees of
// 2*(leaves_per_subtree - 1) hashes each; inside a sub-tree the node at layer l (0 = leaf digests),
// position q sits at  2*(((q >> 1) << (l + 1)) + 2^l - 1) + (q & 1);  sub-tree roots go to cap[t].
// The closed form replaces the reference CUDA's per-access loop (cuda/plonky2_gpu_impl.cuh:315-348).
//
// B200 mapping: one thread per leaf / per node, sponge state in registers, all sub-trees of a layer in
// one launch (the reference launches one 256-thread block per sub-tree, i.e. 16 SMs).  The kernels are
// integer-issue bound (about 2*10^4 instructions per permutation versus 64 B of input), so loads go
// straight through L1 with no shared-memory staging.
#pragma once
#include "poseidon.cuh"

namespace merkle {

using gl::u32;
using gl::u64;

struct TreeShape {
  u64 num_leaves;     // N (power of two)
  u32 log_leaves;     // log2 N
  u32 cap_height;     // c
  u32 sub_log;        // log2(N) - c  : layers below the cap
  u64 sub_digests;    // 2 * (N / 2^c - 1)
};

__host__ __device__ __forceinline__ TreeShape make_shape(u32 log_leaves, u32 cap_height) {
  TreeShape t;
  t.num_leaves = (u64)1 << log_leaves;
  t.log_leaves = log_leaves;
  t.cap_height = cap_height;
  t.sub_log = log_leaves - cap_height;
  t.sub_digests = 2 * (((u64)1 << t.sub_log) - 1);
  return t;
}

// Where node Q of layer l (Q counted across the whole tree, 0 <= Q < N / 2^l) is stored: a pointer into
// `digests` or, for sub-tree roots (l == sub_log), into `cap`.
__device__ __forceinline__ u64* node_slot(const TreeShape& t, u64* digests, u64* cap, u32 l, u64 Q) {
  if (l == t.sub_log) return cap + 4 * Q;
  u32 bits = t.sub_log - l;                    // log2(nodes of this layer per sub-tree)
  u64 tree = Q >> bits;
  u64 q = Q & (((u64)1 << bits) - 1);
  u64 idx = 2 * (((q >> 1) << (l + 1)) + ((u64)1 << l) - 1) + (q & 1);
  return digests + 4 * (tree * t.sub_digests + idx);
}

__device__ __forceinline__ void store_hash(u64* dst, const u64 h[4]) {
  // 32-byte aligned slot -> two 16-byte stores
  reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(h[0], h[1]);
  reinterpret_cast<ulonglong2*>(dst)[1] = make_ulonglong2(h[2], h[3]);
}
__device__ __forceinline__ void load_hash(const u64* src, u64 h[4]) {
  ulonglong2 a = reinterpret_cast<const ulonglong2*>(src)[0];
  ulonglong2 b = reinterpret_cast<const ulonglong2*>(src)[1];
  h[0] = a.x; h[1] = a.y; h[2] = b.x; h[3] = b.y;
}

// hash_or_noop of one leaf whose element j is at base[j * col_stride].  Result canonical.  M selects the exact or the
// optimistic field reduction (gl64.cuh); with the optimistic one the caller must check m.rare and redo exactly.
template <class M>
__device__ __forceinline__ void hash_leaf(const u64* __restrict__ base, u64 col_stride, u32 leaf_len, u64 out[4], M& m) {
  if (leaf_len <= 4) {  // plonk/config.rs:57-63: copy canonical values, zero pad
#pragma unroll
    for (u32 i = 0; i < 4; i++) out[i] = i < leaf_len ? gl::canon(__ldg(base + i * col_stride)) : 0;
    return;
  }
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  // one loop for full and partial chunks (a second inlined copy of the permutation would double the code size);
  // the partial last chunk overwrites only `leaf_len % 8` lanes (hashing.rs:88-91)
  const u64* p = base;
#pragma unroll 1
  for (u32 left = leaf_len; left > 0; left = left > 8 ? left - 8 : 0) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      if ((u32)i < left) s[i] = __ldg(p + i * col_stride);
    p += 8 * col_stride;
    poseidon::permute(s, m);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
}

// One thread per leaf.  leaves element (row, col) at leaves[row * row_stride + col * col_stride].
// leaf_index0: global index of this launch's first leaf (multi-GPU shards / coset blocks hash a sub-range
// of the tree's leaves but write into the whole tree's layout).
#ifndef P2B_HASH_BLOCK
#define P2B_HASH_BLOCK 128
#endif
#ifndef P2B_HASH_MIN_BLOCKS
#define P2B_HASH_MIN_BLOCKS 7
#endif
__global__ void __launch_bounds__(P2B_HASH_BLOCK, P2B_HASH_MIN_BLOCKS)
hash_leaves_kernel(const u64* __restrict__ leaves, u64 row_stride, u64 col_stride, u32 leaf_len, u64 count,
                   u64 leaf_index0, TreeShape shape, u64* __restrict__ digests, u64* __restrict__ cap) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  // tail threads stay alive on a clamped index (they may take part in block-wide barriers inside permute)
  const bool live = i < count;
  if (!live) i = count - 1;
  u64 h[4];
#ifndef P2B_EXACT_ONLY
  gl::Optimistic fast;
  hash_leaf(leaves + i * row_stride, col_stride, leaf_len, h, fast);
  if (fast.rare)  // ~2^-32 per multiplication on random data: redo this leaf with fully repaid reductions
#endif
  {
    gl::Exact exact;
    hash_leaf(leaves + i * row_stride, col_stride, leaf_len, h, exact);
  }
  if (live) store_hash(node_slot(shape, digests, cap, 0, leaf_index0 + i), h);
}

// Progressive sponge for the pipelined (multi-GPU) commit: absorbs columns [col0, col0 + ncols) of every leaf.  col0 is a
// multiple of 8 and ncols a multiple of 8 unless this is the last call (col0 + ncols == leaf_len), so the chunking is
// exactly the one-shot sponge's (hashing.rs:81-104).  Between calls the 12-word state of leaf i lives in
// state[k * count + i]; the last call writes the digest.  leaf_len > 4 (hash_or_noop's copy case never gets here).
template <class M>
__device__ __forceinline__ void absorb_columns(const u64* __restrict__ row, u32 ncols, u64 (&s)[12], M& m) {
  const u64* p = row;
#pragma unroll 1
  for (u32 left = ncols; left > 0; left = left > 8 ? left - 8 : 0) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      if ((u32)i < left) s[i] = __ldg(p + i);
    p += 8;
    poseidon::permute(s, m);
  }
}
__global__ void __launch_bounds__(P2B_HASH_BLOCK, P2B_HASH_MIN_BLOCKS)
absorb_columns_kernel(const u64* __restrict__ leaves, u64 row_stride, u32 col0, u32 ncols, int last, u64 count, u64 leaf_index0,
                      TreeShape shape, u64* __restrict__ state, u64* __restrict__ digests, u64* __restrict__ cap) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < count;
  if (!live) i = count - 1;
  const u64* row = leaves + i * row_stride + col0;
  u64 s[12];
  auto load_state = [&]() {
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = col0 ? state[(u64)k * count + i] : 0;
  };
  load_state();
#ifndef P2B_EXACT_ONLY
  gl::Optimistic fast;
  absorb_columns(row, ncols, s, fast);
  if (fast.rare)  // redo this leaf's chunk exactly from the stored state (not yet overwritten)
#endif
  {
    gl::Exact exact;
    load_state();
    absorb_columns(row, ncols, s, exact);
  }
  if (!live) return;
  if (last) {
    u64 h[4];
#pragma unroll
    for (int k = 0; k < 4; k++) h[k] = gl::canon(s[k]);
    store_hash(node_slot(shape, digests, cap, 0, leaf_index0 + i), h);
  } else {
#pragma unroll
    for (int k = 0; k < 12; k++) state[(u64)k * count + i] = s[k];
  }
}

// One thread per node of layer `l` (l >= 1): parent of nodes 2Q, 2Q+1 of layer l-1.
__global__ void __launch_bounds__(P2B_HASH_BLOCK, P2B_HASH_MIN_BLOCKS)
merkle_layer_kernel(TreeShape shape, u32 l, u64 node0, u64 count, u64* __restrict__ digests, u64* __restrict__ cap) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < count;
  if (!live) i = count - 1;
  u64 Q = node0 + i;
  const u64* left = node_slot(shape, digests, cap, l - 1, 2 * Q);  // siblings are adjacent: left + 4
  u64 a[4], b[4], h[4];
  load_hash(left, a);
  load_hash(left + 4, b);
#ifndef P2B_EXACT_ONLY
  gl::Optimistic fast;
  poseidon::two_to_one(a, b, h, fast);
  if (fast.rare)
#endif
  {
    gl::Exact exact;
    poseidon::two_to_one(a, b, h, exact);
  }
  if (live) store_hash(node_slot(shape, digests, cap, l, Q), h);
}

// Batched permutation (test / micro-benchmark entry): states[i][12] -> permuted, canonical.
__global__ void __launch_bounds__(P2B_HASH_BLOCK, P2B_HASH_MIN_BLOCKS) permute_kernel(u64* __restrict__ states, u64 count, int reps) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < count;
  if (!live) i = count - 1;
  u64 s[12];
#ifndef P2B_EXACT_ONLY
  gl::Optimistic fast;
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = states[i * 12 + k];
  for (int r = 0; r < reps; r++) poseidon::permute(s, fast);
  if (fast.rare)
#endif
  {
    gl::Exact exact;
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[i * 12 + k];
    for (int r = 0; r < reps; r++) poseidon::permute(s, exact);
  }
#pragma unroll
  for (int k = 0; k < 12; k++)
    if (live) states[i * 12 + k] = gl::canon(s[k]);
}

}  // namespace merkle
