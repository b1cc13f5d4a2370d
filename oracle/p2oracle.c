/*
 * p2oracle.c -- TEST INFRASTRUCTURE ONLY (see p2oracle.h).  CPU restatement of the reference's CPU path.
 * Parallelised the way the reference is (rayon): over polynomials for the FFTs
 * (plonky2/src/fri/oracle.rs:717-721, 989-997) and over cap sub-trees + recursive halves for the Merkle
 * tree (plonky2/src/hash/merkle_tree.rs:96-99, 232-243), here with OpenMP.
 */
#include "p2oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "poseidon_tables.h"

typedef unsigned __int128 u128;
#define ORDER P2O_ORDER
#define EPS P2O_EPSILON

static int g_threads = 0;
void p2o_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int p2o_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Field.  field/src/goldilocks_field.rs
 * ---------------------------------------------------------------------------------------------- */

/* to_canonical_u64, goldilocks_field.rs:169-176 */
uint64_t p2o_canon(uint64_t c) { return c >= ORDER ? c - ORDER : c; }

/* Add, goldilocks_field.rs:197-219 (inputs may be non-canonical u64) */
static inline uint64_t f_add(uint64_t a, uint64_t b) {
  uint64_t s = a + b;
  uint64_t over = s < a;
  uint64_t s2 = s + over * EPS;
  if (s2 < s) s2 += EPS; /* double overflow, :205-216 */
  return s2;
}
/* Sub, goldilocks_field.rs:234-256 */
static inline uint64_t f_sub(uint64_t a, uint64_t b) {
  uint64_t d = a - b;
  uint64_t under = a < b;
  uint64_t d2 = d - under * EPS;
  if (d2 > d) d2 -= EPS;
  return d2;
}
/* reduce128, goldilocks_field.rs:345-358 */
static inline uint64_t f_reduce128(u128 x) {
  uint64_t x_lo = (uint64_t)x, x_hi = (uint64_t)(x >> 64);
  uint64_t x_hi_hi = x_hi >> 32, x_hi_lo = x_hi & EPS;
  uint64_t t0 = x_lo - x_hi_hi;
  if (x_lo < x_hi_hi) t0 -= EPS;
  uint64_t t1 = x_hi_lo * EPS;
  uint64_t t2 = t0 + t1; /* add_no_canonicalize_trashing_input, :326-341 */
  if (t2 < t0) t2 += EPS;
  return t2;
}
/* Mul, goldilocks_field.rs:265-272 */
static inline uint64_t f_mul(uint64_t a, uint64_t b) { return f_reduce128((u128)a * (u128)b); }
/* from_noncanonical_u96, goldilocks_field.rs:153-165 ((n_lo, n_hi: u32)) */
static inline uint64_t f_reduce96(uint64_t lo, uint32_t hi) {
  uint64_t t1 = (uint64_t)hi * EPS;
  uint64_t t2 = lo + t1;
  if (t2 < lo) t2 += EPS;
  return t2;
}
/* reduce_u160, plonky2/src/hash/poseidon.rs:40-47 */
static inline uint64_t f_reduce160(u128 n_lo, uint32_t n_hi) {
  uint64_t n_lo_hi = (uint64_t)(n_lo >> 64), n_lo_lo = (uint64_t)n_lo;
  uint64_t reduced_hi = f_reduce96(n_lo_hi, n_hi);
  return f_reduce128(((u128)reduced_hi << 64) + n_lo_lo);
}

uint64_t p2o_add(uint64_t a, uint64_t b) { return p2o_canon(f_add(a, b)); }
uint64_t p2o_sub(uint64_t a, uint64_t b) { return p2o_canon(f_sub(a, b)); }
uint64_t p2o_mul(uint64_t a, uint64_t b) { return p2o_canon(f_mul(a, b)); }

/* exp_u64, field/src/types.rs:347-371 (square and multiply) */
static uint64_t f_exp(uint64_t base, uint64_t e) {
  uint64_t cur = base, prod = 1;
  while (e) {
    if (e & 1) prod = f_mul(prod, cur);
    cur = f_mul(cur, cur);
    e >>= 1;
  }
  return prod;
}
uint64_t p2o_exp(uint64_t base, uint64_t e) { return p2o_canon(f_exp(base, e)); }
/* The reference inverts with a binary-GCD variant (field/src/inversion.rs); the result is the unique
 * field inverse, restated here through Fermat: a^(p-2). */
uint64_t p2o_inverse(uint64_t a) { return p2o_canon(f_exp(a, ORDER - 2)); }
/* inverse_2exp, field/src/types.rs:227-266 */
uint64_t p2o_inverse_2exp(unsigned e) {
  const unsigned T = 32;
  if (e > T) {
    uint64_t inv_t = ORDER - ((ORDER - 1) >> T);
    uint64_t res = inv_t;
    unsigned r = e - T;
    while (r > T) {
      res = f_mul(res, inv_t);
      r -= T;
    }
    return p2o_canon(f_mul(res, ORDER - ((ORDER - 1) >> r)));
  }
  return ORDER - ((ORDER - 1) >> e);
}
/* primitive_root_of_unity, field/src/types.rs:268-272; POWER_OF_TWO_GENERATOR goldilocks_field.rs:89 */
uint64_t p2o_primitive_root_of_unity(unsigned n_log) {
  uint64_t b = 1753635133440165772ull;
  for (unsigned i = 0; i < 32 - n_log; i++) b = f_mul(b, b);
  return p2o_canon(b);
}
/* coset_shift = MULTIPLICATIVE_GROUP_GENERATOR = 7, goldilocks_field.rs:82, types.rs:431-433 */
uint64_t p2o_coset_shift(void) { return 7; }

/* ------------------------------------------------------------------------------------------------
 * Bit reversal.  plonky2/src/util/mod.rs:55-63, util/src/lib.rs:188-237
 * ---------------------------------------------------------------------------------------------- */
uint64_t p2o_reverse_bits(uint64_t n, unsigned num_bits) {
  uint64_t r = 0;
  for (unsigned i = 0; i < num_bits; i++) r |= ((n >> i) & 1) << (num_bits - 1 - i);
  return r;
}
static unsigned log2_strict(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
void p2o_reverse_index_bits_in_place(uint64_t* v, size_t n) {
  unsigned lb = log2_strict(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = p2o_reverse_bits(i, lb);
    if (i < j) {
      uint64_t t = v[i];
      v[i] = v[j];
      v[j] = t;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * FFT.  field/src/fft.rs
 * ---------------------------------------------------------------------------------------------- */
/* fft_root_table, fft.rs:15-34: row lg_m (1..=lg_n) = powers of root(lg_m), length max(2^(lg_m-1), 2). */
size_t p2o_root_table_len(unsigned n_log) {
  size_t t = 0;
  for (unsigned lg_m = 1; lg_m <= n_log; lg_m++) {
    size_t h = (size_t)1 << (lg_m - 1);
    t += h < 2 ? 2 : h;
  }
  return t;
}
void p2o_fft_root_table_concat(unsigned n_log, uint64_t* out) {
  size_t off = 0;
  for (unsigned lg_m = 1; lg_m <= n_log; lg_m++) {
    size_t h = (size_t)1 << (lg_m - 1);
    size_t len = h < 2 ? 2 : h;
    uint64_t base = p2o_primitive_root_of_unity(lg_m), cur = 1;
    for (size_t i = 0; i < len; i++) {
      out[off + i] = p2o_canon(cur);
      cur = f_mul(cur, base);
    }
    off += len;
  }
}

/* per-size cached twiddle rows (row s holds omega_{2^(s+1)}^j, j < 2^s) */
static uint64_t* g_tw[34];
static void ensure_twiddles(unsigned n_log) {
#pragma omp critical(p2o_tw)
  {
    for (unsigned s = 0; s < n_log; s++) {
      if (!g_tw[s]) {
        size_t h = (size_t)1 << s;
        uint64_t* row = (uint64_t*)malloc(h * sizeof(uint64_t));
        uint64_t base = p2o_primitive_root_of_unity(s + 1), cur = 1;
        for (size_t i = 0; i < h; i++) {
          row[i] = cur;
          cur = f_mul(cur, base);
        }
        g_tw[s] = row;
      }
    }
  }
}

/* fft_classic, fft.rs:188-229 + fft_classic_simd :107-180 (scalar shape): bit-reverse, replicate for the
 * zero tail (r), then radix-2 DIT stages lg_half_m = r..lg_n. */
void p2o_fft(uint64_t* v, unsigned n_log, unsigned r) {
  size_t n = (size_t)1 << n_log;
  ensure_twiddles(n_log);
  p2o_reverse_index_bits_in_place(v, n);
  if (r > 0) {
    size_t mask = ~(((size_t)1 << r) - 1);
    for (size_t i = 0; i < n; i++) v[i] = v[i & mask];
  }
  for (unsigned lg_half_m = r; lg_half_m < n_log; lg_half_m++) {
    size_t half_m = (size_t)1 << lg_half_m, m = half_m << 1;
    const uint64_t* tw = g_tw[lg_half_m];
    for (size_t k = 0; k < n; k += m) {
      for (size_t j = 0; j < half_m; j++) {
        uint64_t t = f_mul(tw[j], v[k + half_m + j]);
        uint64_t u = v[k + j];
        v[k + j] = f_add(u, t);
        v[k + half_m + j] = f_sub(u, t);
      }
    }
  }
  for (size_t i = 0; i < n; i++) v[i] = p2o_canon(v[i]);
}

/* ifft_with_options, fft.rs:73-103 */
void p2o_ifft(uint64_t* b, unsigned n_log) {
  size_t n = (size_t)1 << n_log;
  uint64_t n_inv = p2o_inverse_2exp(n_log);
  p2o_fft(b, n_log, 0);
  b[0] = p2o_mul(b[0], n_inv);
  if (n > 1) b[n / 2] = p2o_mul(b[n / 2], n_inv);
  for (size_t i = 1; i < n / 2; i++) {
    size_t j = n - i;
    uint64_t ci = p2o_mul(b[j], n_inv), cj = p2o_mul(b[i], n_inv);
    b[i] = ci;
    b[j] = cj;
  }
}

/* coset_fft_with_options, field/src/polynomial/mod.rs:286-299 */
void p2o_coset_fft(uint64_t* c, unsigned n_log, uint64_t shift, unsigned r) {
  size_t n = (size_t)1 << n_log;
  uint64_t pw = 1;
  for (size_t i = 0; i < n; i++) {
    c[i] = f_mul(pw, c[i]);
    pw = f_mul(pw, shift);
  }
  p2o_fft(c, n_log, r);
}
/* coset_ifft, field/src/polynomial/mod.rs:64-77 */
void p2o_coset_ifft(uint64_t* v, unsigned n_log, uint64_t shift) {
  size_t n = (size_t)1 << n_log;
  p2o_ifft(v, n_log);
  uint64_t sinv = p2o_inverse(shift), pw = 1;
  for (size_t i = 0; i < n; i++) {
    v[i] = p2o_mul(v[i], pw);
    pw = f_mul(pw, sinv);
  }
}
/* p.lde(rate_bits).coset_fft_with_options(coset_shift, Some(rate_bits)), fri/oracle.rs:992-995 */
void p2o_lde_coset_fft(const uint64_t* coeffs, unsigned n_log, unsigned rate_bits, uint64_t* out) {
  size_t n = (size_t)1 << n_log, N = n << rate_bits;
  memcpy(out, coeffs, n * sizeof(uint64_t));
  memset(out + n, 0, (N - n) * sizeof(uint64_t));
  p2o_coset_fft(out, n_log + rate_bits, p2o_coset_shift(), rate_bits);
}
/* evaluate_naive / eval on the subgroup, fft.rs:257-276 and polynomial/mod.rs:483-501 */
void p2o_naive_coset_eval(const uint64_t* coeffs, size_t ncoeffs, unsigned n_log, uint64_t shift,
                          uint64_t* out) {
  size_t n = (size_t)1 << n_log;
  uint64_t w = p2o_primitive_root_of_unity(n_log);
  uint64_t x = shift;
#pragma omp parallel for schedule(static)
  for (size_t m = 0; m < n; m++) {
    uint64_t xm = f_mul(shift, f_exp(w, m));
    uint64_t acc = 0;
    for (size_t i = ncoeffs; i-- > 0;) acc = f_add(f_mul(acc, xm), coeffs[i]);
    out[m] = p2o_canon(acc);
  }
  (void)x;
}

/* ------------------------------------------------------------------------------------------------
 * Poseidon.  plonky2/src/hash/poseidon.rs
 * ---------------------------------------------------------------------------------------------- */
#define W 12
/* constant_layer, poseidon.rs:482-493 */
static inline void constant_layer(uint64_t* s, int round_ctr) {
  for (int i = 0; i < W; i++) s[i] = f_add(s[i], P2_ROUND_CONSTANTS[i + W * round_ctr]);
}
/* sbox_monomial, poseidon.rs:525-532 */
static inline uint64_t sbox(uint64_t x) {
  uint64_t x2 = f_mul(x, x), x4 = f_mul(x2, x2), x3 = f_mul(x, x2);
  return f_mul(x3, x4);
}
/* mds_row_shf + mds_layer, poseidon.rs:172-194, 236-260 */
static inline void mds_layer(uint64_t* s) {
  uint64_t out[W], d[2 * W];
  memcpy(d, s, W * sizeof(uint64_t));
  memcpy(d + W, s, W * sizeof(uint64_t)); /* doubled copy: s[(i + r) % W] == d[i + r] */
#pragma GCC unroll 12
  for (int r = 0; r < W; r++) {
    u128 acc = 0;
#pragma GCC unroll 12
    for (int i = 0; i < W; i++) acc += (u128)d[i + r] * (u128)P2_MDS_CIRC[i];
    acc += (u128)s[r] * (u128)P2_MDS_DIAG[r];
    out[r] = f_reduce96((uint64_t)acc, (uint32_t)(acc >> 64));
  }
  memcpy(s, out, sizeof(out));
}
/* full_rounds, poseidon.rs:560-572 */
static inline void full_rounds(uint64_t* s, int* round_ctr) {
  for (int k = 0; k < 4; k++) {
    constant_layer(s, *round_ctr);
    for (int i = 0; i < W; i++) s[i] = sbox(s[i]);
    mds_layer(s);
    (*round_ctr)++;
  }
}
/* partial_rounds (fast), poseidon.rs:574-588 with helpers :310-365 (init), :398-427 (fast layer) */
static inline void partial_rounds(uint64_t* s, int* round_ctr) {
  for (int i = 0; i < W; i++) s[i] = f_add(s[i], P2_PARTIAL_FIRST_RC[i]);
  {
    uint64_t res[W];
    memset(res, 0, sizeof(res));
    res[0] = s[0];
    /* same sums as the reference's loop nest (:319-333), accumulated per output column in a u160 so that the
     * CPU baseline is not penalised by 121 separate reductions (field addition is exact, the result is equal) */
    for (int c = 1; c < W; c++) {
      u128 lo = 0;
      uint32_t hi = 0;
#pragma GCC unroll 11
      for (int r = 1; r < W; r++) {
        u128 t = (u128)s[r] * (u128)P2_PARTIAL_INIT_MATRIX[(r - 1) * 11 + (c - 1)];
        u128 nl = lo + t;
        hi += nl < lo;
        lo = nl;
      }
      res[c] = f_reduce160(lo, hi);
    }
    memcpy(s, res, sizeof(res));
  }
  for (int r = 0; r < 22; r++) {
    s[0] = sbox(s[0]);
    s[0] = f_add(s[0], P2_PARTIAL_RC[r]);
    /* mds_partial_layer_fast: d = M00*s0 + sum w_hat[i-1]*s[i] in a u160 accumulator */
    u128 lo = 0;
    uint32_t hi = 0;
#pragma GCC unroll 11
    for (int i = 1; i < W; i++) {
      u128 t = (u128)s[i] * (u128)P2_PARTIAL_W_HATS[r * 11 + i - 1];
      u128 nl = lo + t;
      hi += nl < lo;
      lo = nl;
    }
    {
      u128 t = (u128)s[0] * (u128)(P2_MDS_CIRC[0] + P2_MDS_DIAG[0]);
      u128 nl = lo + t;
      hi += nl < lo;
      lo = nl;
    }
    uint64_t d = f_reduce160(lo, hi);
    uint64_t s0 = s[0];
    s[0] = d;
#pragma GCC unroll 11
    for (int i = 1; i < W; i++) /* multiply_accumulate, goldilocks_field.rs:123-127 */
      s[i] = f_reduce128((u128)s[i] + (u128)s0 * (u128)P2_PARTIAL_VS[r * 11 + i - 1]);
  }
  *round_ctr += 22;
}
/* poseidon, poseidon.rs:590-606 */
void p2o_poseidon(uint64_t s[12]) {
  int rc = 0;
  full_rounds(s, &rc);
  partial_rounds(s, &rc);
  full_rounds(s, &rc);
  for (int i = 0; i < W; i++) s[i] = p2o_canon(s[i]);
}
/* poseidon_naive, poseidon.rs:608-630 */
void p2o_poseidon_naive(uint64_t s[12]) {
  int rc = 0;
  full_rounds(s, &rc);
  for (int k = 0; k < 22; k++) {
    constant_layer(s, rc);
    s[0] = sbox(s[0]);
    mds_layer(s);
    rc++;
  }
  full_rounds(s, &rc);
  for (int i = 0; i < W; i++) s[i] = p2o_canon(s[i]);
}

/* hash_n_to_m_no_pad (num_outputs = 4), plonky2/src/hash/hashing.rs:81-104: overwrite-mode sponge */
void p2o_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]) {
  uint64_t st[W];
  memset(st, 0, sizeof(st));
  for (size_t off = 0; off < len; off += 8) {
    size_t c = len - off < 8 ? len - off : 8;
    for (size_t i = 0; i < c; i++) st[i] = in[off + i];
    p2o_poseidon(st);
  }
  /* len == 0: no absorb, squeeze from the zero state (the `loop` at :96 pushes state[0..4] directly) */
  for (int i = 0; i < 4; i++) out[i] = p2o_canon(st[i]);
}
/* hash_or_noop, plonky2/src/plonk/config.rs:56-67 */
void p2o_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]) {
  if (len <= 4) {
    for (size_t i = 0; i < 4; i++) out[i] = i < len ? p2o_canon(in[i]) : 0;
  } else {
    p2o_hash_no_pad(in, len, out);
  }
}
/* compress, hashing.rs:65-72 */
void p2o_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t st[W];
  memset(st, 0, sizeof(st));
  memcpy(st, l, 32);
  memcpy(st + 4, r, 32);
  p2o_poseidon(st);
  memcpy(out, st, 32);
}

/* ------------------------------------------------------------------------------------------------
 * Merkle tree.  plonky2/src/hash/merkle_tree.rs
 * ---------------------------------------------------------------------------------------------- */
/* fill_subtree, merkle_tree.rs:78-105.  digests_buf: dlen hashes; leaves: (dlen/2+1) rows. */
static void fill_subtree(uint64_t* digests_buf, size_t dlen, const uint64_t* leaves, size_t nleaves,
                         size_t leaf_len, uint64_t out[4], int depth) {
  if (dlen == 0) {
    p2o_hash_or_noop(leaves, leaf_len, out);
    return;
  }
  size_t half = dlen / 2;
  uint64_t* left_buf = digests_buf;               /* [0, half-1) recursive, [half-1] left digest */
  uint64_t* left_digest_mem = digests_buf + 4 * (half - 1);
  uint64_t* right_digest_mem = digests_buf + 4 * half;
  uint64_t* right_buf = digests_buf + 4 * (half + 1);
  uint64_t ld[4], rd[4];
  if (depth > 0) {
#pragma omp task shared(ld) firstprivate(left_buf, half, leaves, nleaves, leaf_len, depth)
    fill_subtree(left_buf, half - 1, leaves, nleaves / 2, leaf_len, ld, depth - 1);
#pragma omp task shared(rd) firstprivate(right_buf, half, leaves, nleaves, leaf_len, depth)
    fill_subtree(right_buf, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len, rd, depth - 1);
#pragma omp taskwait
  } else {
    fill_subtree(left_buf, half - 1, leaves, nleaves / 2, leaf_len, ld, 0);
    fill_subtree(right_buf, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len, rd, 0);
  }
  memcpy(left_digest_mem, ld, 32);
  memcpy(right_digest_mem, rd, 32);
  p2o_two_to_one(ld, rd, out);
}

/* MerkleTree::new + fill_digests_buf, merkle_tree.rs:283-319, 210-244 */
int p2o_merkle_tree(const uint64_t* leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    uint64_t* digests, uint64_t* cap) {
  unsigned lg = log2_strict(num_leaves);
  if (((size_t)1 << lg) != num_leaves || cap_height > lg) return -1;
  size_t ncap = (size_t)1 << cap_height;
  size_t num_digests = 2 * (num_leaves - ncap);
  if (num_digests == 0) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < num_leaves; i++) p2o_hash_or_noop(leaves + i * leaf_len, leaf_len, cap + 4 * i);
    return 0;
  }
  size_t sub_d = num_digests >> cap_height, sub_l = num_leaves >> cap_height;
  /* one task per sub-tree; recursive halves split further (rayon::join) until ~8x threads tasks exist */
  int depth = 0;
  {
    size_t tasks = ncap;
    size_t want = (size_t)p2o_get_threads() * 8;
    while (tasks < want && ((size_t)1 << (depth + 1)) <= sub_l / 2) {
      tasks <<= 1;
      depth++;
    }
  }
#pragma omp parallel
#pragma omp single
  {
    for (size_t t = 0; t < ncap; t++) {
#pragma omp task firstprivate(t)
      fill_subtree(digests + 4 * t * sub_d, sub_d, leaves + t * sub_l * leaf_len, sub_l, leaf_len, cap + 4 * t,
                   depth);
    }
#pragma omp taskwait
  }
  return 0;
}

/* MerkleTree::prove, merkle_tree.rs:392-440 */
void p2o_merkle_prove(const uint64_t* digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      uint64_t* siblings) {
  unsigned num_layers = log2_strict(num_leaves) - cap_height;
  size_t ncap = (size_t)1 << cap_height;
  size_t tree_len = (2 * (num_leaves - ncap)) >> cap_height;
  size_t tree_index = leaf_index >> num_layers;
  const uint64_t* tree = digests + 4 * tree_len * tree_index;
  size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
  for (unsigned i = 0; i < num_layers; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    size_t sibling_index = 2 * siblings_index + (1 - parity);
    memcpy(siblings + 4 * i, tree + 4 * sibling_index, 32);
  }
}
/* verify_merkle_proof_to_cap, merkle_proofs.rs:53-81 */
int p2o_merkle_verify(const uint64_t* leaf, size_t leaf_len, size_t leaf_index, const uint64_t* cap,
                      unsigned cap_height, const uint64_t* siblings, size_t num_siblings) {
  (void)cap_height;
  uint64_t cur[4], nxt[4];
  size_t index = leaf_index;
  p2o_hash_or_noop(leaf, leaf_len, cur);
  for (size_t i = 0; i < num_siblings; i++) {
    size_t bit = index & 1;
    index >>= 1;
    if (bit)
      p2o_two_to_one(siblings + 4 * i, cur, nxt);
    else
      p2o_two_to_one(cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, 32);
  }
  return memcmp(cur, cap + 4 * index, 32) == 0;
}

/* ------------------------------------------------------------------------------------------------
 * PolynomialBatch.  plonky2/src/fri/oracle.rs
 * ---------------------------------------------------------------------------------------------- */
/* from_coeffs, oracle.rs:911-977: lde_values (:979-1004) -> transpose (util/mod.rs:23-53) ->
 * reverse_index_bits_in_place -> MerkleTree::new */
int p2o_batch_from_coeffs(const uint64_t* coeffs, unsigned n_log, size_t P, unsigned rate_bits,
                          unsigned cap_height, const uint64_t* salt, uint64_t* leaves_out,
                          uint64_t* digests_out, uint64_t* cap_out) {
  size_t n = (size_t)1 << n_log, N = n << rate_bits;
  unsigned N_log = n_log + rate_bits;
  size_t salt_size = salt ? 4 : 0, leaf_len = P + salt_size;
  if (cap_height > N_log) return -1;
  ensure_twiddles(N_log);
  uint64_t* leaves = leaves_out ? leaves_out : (uint64_t*)malloc(N * leaf_len * sizeof(uint64_t));
  if (!leaves) return -2;
  /* "FFT + blinding": one task per polynomial (par_iter), then transpose + bit-reverse fused into the
   * scatter: leaf[L][col] = lde[col][reverse_bits(L)] */
#pragma omp parallel
  {
    uint64_t* buf = (uint64_t*)malloc(N * sizeof(uint64_t));
#pragma omp for schedule(dynamic, 1)
    for (size_t c = 0; c < leaf_len; c++) {
      if (c < P)
        p2o_lde_coset_fft(coeffs + c * n, n_log, rate_bits, buf);
      else
        for (size_t i = 0; i < N; i++) buf[i] = p2o_canon(salt[(c - P) * N + i]);
      for (size_t L = 0; L < N; L++) leaves[L * leaf_len + c] = buf[p2o_reverse_bits(L, N_log)];
    }
    free(buf);
  }
  int rc = 0;
  if (digests_out || cap_out) {
    size_t ncap = (size_t)1 << cap_height;
    size_t nd = 2 * (N - ncap);
    uint64_t* dg = digests_out ? digests_out : (uint64_t*)malloc((nd ? nd : 1) * 32);
    uint64_t* cp = cap_out ? cap_out : (uint64_t*)malloc(ncap * 32);
    rc = p2o_merkle_tree(leaves, N, leaf_len, cap_height, dg, cp);
    if (!digests_out) free(dg);
    if (!cap_out) free(cp);
  }
  if (!leaves_out) free(leaves);
  return rc;
}

/* from_values, oracle.rs:709-731: ifft per polynomial then from_coeffs */
int p2o_batch_from_values(const uint64_t* values, unsigned n_log, size_t P, unsigned rate_bits,
                          unsigned cap_height, const uint64_t* salt, uint64_t* coeffs_out,
                          uint64_t* leaves_out, uint64_t* digests_out, uint64_t* cap_out) {
  size_t n = (size_t)1 << n_log;
  uint64_t* coeffs = coeffs_out ? coeffs_out : (uint64_t*)malloc(P * n * sizeof(uint64_t));
  if (!coeffs) return -2;
  ensure_twiddles(n_log);
  memcpy(coeffs, values, P * n * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < P; c++) p2o_ifft(coeffs + c * n, n_log);
  int rc = p2o_batch_from_coeffs(coeffs, n_log, P, rate_bits, cap_height, salt, leaves_out, digests_out, cap_out);
  if (!coeffs_out) free(coeffs);
  return rc;
}
