/*
 * p2oracle -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the reference's (sideprotocol/plonky2-gpu) CPU algorithm for the
 * polynomial-commitment hot path: Goldilocks field, FFT/iFFT, coset LDE, bit-reversal, Poseidon sponge,
 * MerkleTree::new and PolynomialBatch::from_values/from_coeffs (compute_quotient_polys, the permutation argument and the
 * FRI opening proof are restated in Python: oracle/quotient.py, oracle/recursion_gates.py, oracle/fri.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library.  The product (plonky2-gpu_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference's Rust CPU path cannot be built in this image (no cargo/rustc), so the
 * oracle is pinned by (i) every known-answer vector the reference's own tests hold for this path -- the 4
 * Poseidon permutation vectors (plonky2/src/hash/poseidon_goldilocks.rs:289-310), the n=256 bit-reversal
 * table (plonky2/src/util/mod.rs:82-102), inverse_2exp(18)/(21) literals (cuda/test.cu:195,
 * cuda/plonky2_gpu.cu:746), the partial-products example (plonky2/src/util/partial_products.rs:115-140)
 * -- (ii) the reference's property tests restated (FFT == naive evaluation, coset FFT == naive, every
 * Merkle proof verifies, fast == naive partial rounds) and (iii) on the GPU box, the reference's own CUDA
 * kernels compiled unmodified from /root/reference/cuda into oracle/_ref/ and run on the same inputs
 * (tests/test_ref_cuda_crosscheck.py).  LDE matrices / caps / quotient values have no stored vectors in
 * the reference; for those the parity is "pinned transitively + by the reference CUDA run".
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#ifndef P2ORACLE_H
#define P2ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2O_ORDER 0xFFFFFFFF00000001ull
#define P2O_EPSILON 0xFFFFFFFFull

/* ---- field (field/src/goldilocks_field.rs, field/src/types.rs) ---- */
uint64_t p2o_add(uint64_t a, uint64_t b);
uint64_t p2o_sub(uint64_t a, uint64_t b);
uint64_t p2o_mul(uint64_t a, uint64_t b);
uint64_t p2o_canon(uint64_t a);
uint64_t p2o_exp(uint64_t base, uint64_t e);
uint64_t p2o_inverse(uint64_t a);
uint64_t p2o_inverse_2exp(unsigned e);
uint64_t p2o_primitive_root_of_unity(unsigned n_log);
uint64_t p2o_coset_shift(void);

/* ---- bit reversal (util/src/lib.rs:188-237, plonky2/src/util/mod.rs:55-63) ---- */
uint64_t p2o_reverse_bits(uint64_t n, unsigned num_bits);
void p2o_reverse_index_bits_in_place(uint64_t* v, size_t n);

/* ---- FFT (field/src/fft.rs) ---- */
/* Concatenated root table exactly as `fft_root_table(n).concat()` (circuit_builder.rs:850-851): row lg_m=1
 * has 2 entries, row lg_m has max(2^(lg_m-1),2) entries.  out must hold p2o_root_table_len(n_log). */
size_t p2o_root_table_len(unsigned n_log);
void p2o_fft_root_table_concat(unsigned n_log, uint64_t* out);
void p2o_fft(uint64_t* values, unsigned n_log, unsigned zero_factor_r); /* in place, natural order   */
void p2o_ifft(uint64_t* values, unsigned n_log);                        /* in place, natural order   */
void p2o_coset_fft(uint64_t* coeffs, unsigned n_log, uint64_t shift, unsigned zero_factor_r);
void p2o_coset_ifft(uint64_t* values, unsigned n_log, uint64_t shift);
/* out[N] = lde(coeffs, rate_bits).coset_fft(g=7, zero_factor = rate_bits) (polynomial/mod.rs:205-299) */
void p2o_lde_coset_fft(const uint64_t* coeffs, unsigned n_log, unsigned rate_bits, uint64_t* out);
/* naive O(n^2) evaluation on shift*<omega_n> for the property tests (fft.rs:243-276) */
void p2o_naive_coset_eval(const uint64_t* coeffs, size_t ncoeffs, unsigned n_log, uint64_t shift, uint64_t* out);

/* ---- Poseidon (plonky2/src/hash/poseidon.rs, hashing.rs, plonk/config.rs) ---- */
void p2o_poseidon(uint64_t state[12]);
void p2o_poseidon_naive(uint64_t state[12]);
void p2o_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]);
void p2o_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]);
void p2o_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

/* ---- Merkle tree (plonky2/src/hash/merkle_tree.rs, merkle_proofs.rs) ---- */
/* leaves: row-major [num_leaves][leaf_len]; digests: 2*(num_leaves - 2^cap_height) hashes of 4 u64 in the
 * reference's recursive layout; cap: 2^cap_height hashes.  Returns 0, or -1 if cap_height > log2(leaves). */
int p2o_merkle_tree(const uint64_t* leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    uint64_t* digests, uint64_t* cap);
/* MerkleTree::prove (merkle_tree.rs:392-440): writes (log2(num_leaves) - cap_height) siblings. */
void p2o_merkle_prove(const uint64_t* digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      uint64_t* siblings);
/* verify_merkle_proof_to_cap (merkle_proofs.rs:53-81): 1 if ok. */
int p2o_merkle_verify(const uint64_t* leaf, size_t leaf_len, size_t leaf_index, const uint64_t* cap,
                      unsigned cap_height, const uint64_t* siblings, size_t num_siblings);

/* ---- PolynomialBatch (plonky2/src/fri/oracle.rs:709-731, 911-1018) ---- */
/* values / coeffs: column-major [P][n].  salt: NULL or column-major [4][N] (blinding columns appended to
 * every leaf, oracle.rs:998-1002 -- the reference draws them at random; parity needs them as an input).
 * Outputs (any may be NULL): coeffs_out [P][n]; leaves_out row-major [N][P+salt] in the reference's
 * bit-reversed leaf order; digests_out; cap_out. */
int p2o_batch_from_values(const uint64_t* values, unsigned n_log, size_t P, unsigned rate_bits,
                          unsigned cap_height, const uint64_t* salt, uint64_t* coeffs_out,
                          uint64_t* leaves_out, uint64_t* digests_out, uint64_t* cap_out);
int p2o_batch_from_coeffs(const uint64_t* coeffs, unsigned n_log, size_t P, unsigned rate_bits,
                          unsigned cap_height, const uint64_t* salt, uint64_t* leaves_out,
                          uint64_t* digests_out, uint64_t* cap_out);

void p2o_set_threads(int n);
int p2o_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
