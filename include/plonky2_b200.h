/*
 * plonky2_b200.h -- C ABI of the B200-native polynomial-commitment library (libplonky2_b200.so).
 *
 * This is the drop-in boundary for the path the reference accelerates through its `plonky2_cuda` crate:
 *   reference FFI declarations   cuda/src/lib.rs:52-145
 *   reference definitions        cuda/plonky2_gpu.cu:57-785
 *   reference callers            plonky2/src/fri/oracle.rs:279-545 (from_values_with_gpu), :547-700
 *                                (from_coeffs_with_gpu), plonky2/src/plonk/prover.rs:535-568 (my_prove)
 *
 * Two layers are exported:
 *   1. A generic API (`p2b_*`): 64-bit sizes, explicit host/device flags, int status codes with
 *      p2b_last_error().  It implements PolynomialBatch::from_values / from_coeffs
 *      (plonky2/src/fri/oracle.rs:709-731, 911-977), MerkleTree::new / prove
 *      (plonky2/src/hash/merkle_tree.rs:283-319, 392-440), get_lde_values (oracle.rs:1007-1018) and
 *      compute_quotient_polys (plonky2/src/plonk/prover.rs:790-1034) for any (rows, columns, rate_bits,
 *      cap_height, gate set).
 *   2. The reference's own six symbols (`init`, `ifft`, `build_merkle_tree`, `merkle_tree_from_values`,
 *      `merkle_tree_from_coeffs`, `compute_quotient_polys`) plus `transpose` and `fft_blinding`, with the
 *      reference's exact signatures, in-place device-memory layout and by-value error struct, so
 *      cuda/src/lib.rs links against this library unchanged.  See INTEGRATION.md.
 *
 * All data are Goldilocks field elements stored as little-endian u64.  Everything the library returns is
 * canonical (< p).  There is no CPU fallback: every entry fails with P2B_ERR_CUDA if no sm_100 device is
 * usable.
 */
#ifndef PLONKY2_B200_H
#define PLONKY2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------------
 * Status codes
 * ------------------------------------------------------------------------------------------------- */
enum {
  P2B_OK = 0,
  P2B_ERR_INVALID = 1, /* bad argument (e.g. cap_height > log2(leaves), merkle_tree.rs:285-290)          */
  P2B_ERR_CUDA = 2,    /* a CUDA call or kernel launch failed; p2b_last_error() has the CUDA string     */
  P2B_ERR_OOM = 3,     /* device allocation failed                                                      */
  P2B_ERR_UNSUPPORTED = 4
};

typedef struct p2b_ctx p2b_ctx;     /* per-device context: streams, twiddle tables, workspace           */
typedef struct p2b_batch p2b_batch; /* a committed PolynomialBatch resident on the device               */

/* Thread-local message of the last failing call ("" if none). */
const char* p2b_last_error(void);
/* Library / build identification, e.g. "plonky2_b200 0.1 sm_100a". */
const char* p2b_version(void);

/* ---------------------------------------------------------------------------------------------------
 * Context.  Replaces the reference's externally constructed `CudaInvContext`
 * (plonky2/src/fri/oracle.rs:75-109: streams + root tables + shift powers + one big cache buffer).
 * ------------------------------------------------------------------------------------------------- */
int p2b_ctx_create(int device /* -1 = current */, p2b_ctx** out);
/* Stage tracing after the reference's TimingTree (plonky2/src/util/timing.rs:8-192): every stage the library enqueues carries
 * the name of the reference's `timed!` scope ("IFFT", "FFT + blinding", "build Merkle tree: ...", fri/oracle.rs:717-966;
 * "compute quotient polys", "compute partial products", "construct the opening set", "compute opening proofs",
 * plonk/prover.rs:112-236) as an NVTX range (visible to nsys / ncu) and, once enabled here, as a CUDA-event pair.
 * p2b_ctx_trace_report synchronises and writes "name\tcalls\ttotal_ms\n" per stage, then clears the record. */
int p2b_ctx_trace(p2b_ctx* ctx, int enable);
int p2b_ctx_trace_report(p2b_ctx* ctx, char* buf, uint64_t buf_len);
void p2b_ctx_destroy(p2b_ctx* ctx);
/* The CUDA stream (cudaStream_t) the context launches on; callers may enqueue their own copies on it. */
void* p2b_ctx_stream(p2b_ctx* ctx);
int p2b_ctx_synchronize(p2b_ctx* ctx);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
uint64_t p2b_ctx_launch_count(const p2b_ctx* ctx);
/* Live timing of the dominant kernel (leaf hashing) with CUDA events on the stream it is launched on: enable, run
 * commits, then read the summed duration and the number of launches timed (bench.py's roofline figure). */
int p2b_ctx_time_leaf_hash(p2b_ctx* ctx, int enable);
/* Test hook: the NTT kernels use an optimistic field reduction and redo a tile exactly when it reports its rare case
 * (probability ~2^-32 per butterfly); enabling this makes every tile take that redo path. */
int p2b_ctx_debug_force_exact_redo(p2b_ctx* ctx, int enable);
int p2b_ctx_leaf_hash_time(p2b_ctx* ctx, double* total_ms, uint64_t* launches);

/* ---------------------------------------------------------------------------------------------------
 * PolynomialBatch::from_values / from_coeffs
 *   values / coeffs : column-major [P][n], n = 2^n_log (one polynomial after another, as the reference
 *                     flattens them: fri/oracle.rs:352-362).  Host or device pointer per `on_host`.
 *   salt            : NULL (blinding = false) or column-major [4][n << rate_bits] blinding columns in
 *                     LDE-domain order (the reference draws them with F::rand_vec, oracle.rs:998-1002;
 *                     the caller supplies them so results are reproducible).  SALT_SIZE = 4, oracle.rs:41.
 * The batch keeps, on the device: coefficients [P][n]; leaves row-major [N][P + salt] in the reference's
 * bit-reversed leaf order; digests (2*(N - 2^cap_height) x 4, reference layout); cap (2^cap_height x 4).
 * Host buffer lifetime: the calls enqueue their uploads and return without waiting for them.  Pageable host memory is
 * staged by the runtime before the call returns; PINNED host inputs (values, coeffs, salt) are read by the copy engine
 * later and must stay valid and unmodified until the context has been synchronised (p2b_ctx_synchronize, or any
 * p2b_batch_get_* on the returned batch).
 * ------------------------------------------------------------------------------------------------- */
#define P2B_SALT_SIZE 4

int p2b_commit_from_values(p2b_ctx* ctx, const uint64_t* values, int values_on_host, uint32_t n_log, uint64_t P,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, int salt_on_host,
                           p2b_batch** out);
int p2b_commit_from_coeffs(p2b_ctx* ctx, const uint64_t* coeffs, int coeffs_on_host, uint32_t n_log, uint64_t P,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, int salt_on_host,
                           p2b_batch** out);
/* from_values with the host-side data movement of the reference's caller folded in (fri/oracle.rs:352-362, 403-407):
 * host values are uploaded in column groups while earlier groups are already being transformed, and, if
 * coeffs_host_out (pinned host memory, [P][n]) is given, the coefficients are copied back on the D2H engine while the
 * LDE and the Merkle tree are computed.  The copy is complete once the context stream has been synchronised
 * (p2b_ctx_synchronize, or any p2b_batch_get_*). */
int p2b_commit_from_values_ex(p2b_ctx* ctx, const uint64_t* values, int values_on_host, uint32_t n_log, uint64_t P,
                              uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, int salt_on_host,
                              uint64_t* coeffs_host_out, p2b_batch** out);
void p2b_batch_destroy(p2b_batch* b);

typedef struct {
  uint32_t degree_log; /* n_log                                    (PolynomialBatch.degree_log) */
  uint32_t rate_bits;
  uint32_t cap_height;
  uint32_t salt_size;  /* 0 or 4                                   (blinding)                   */
  uint64_t num_polys;  /* P                                                                     */
  uint64_t num_leaves; /* N = n << rate_bits                                                    */
  uint64_t leaf_len;   /* P + salt_size                                                         */
  uint64_t num_digests;/* 2 * (N - 2^cap_height)                                                */
} p2b_batch_info;
int p2b_batch_get_info(const p2b_batch* b, p2b_batch_info* out);

/* Device pointers of the resident arrays (any out pointer may be NULL). */
int p2b_batch_device_ptrs(const p2b_batch* b, uint64_t** coeffs, uint64_t** leaves, uint64_t** digests,
                          uint64_t** cap);

/* Copies to HOST buffers (synchronous on the context stream). */
int p2b_batch_get_coeffs(const p2b_batch* b, uint64_t* out /* [P][n] */);
int p2b_batch_get_cap(const p2b_batch* b, uint64_t* out /* [2^cap_height][4] */);
int p2b_batch_get_digests(const p2b_batch* b, uint64_t* out /* [num_digests][4] */);
int p2b_batch_get_leaves(const p2b_batch* b, uint64_t first_leaf, uint64_t count, uint64_t* out /* [count][leaf_len] */);
/* get_lde_values(index, step) (fri/oracle.rs:1007-1018): the row at reverse_bits(index*step), salt stripped. */
int p2b_batch_get_lde_values(const p2b_batch* b, uint64_t index, uint64_t step, uint64_t* out /* [P] */);
/* MerkleTree::prove (merkle_tree.rs:392-440): (log2 N - cap_height) sibling hashes for each leaf index. */
int p2b_batch_prove(const p2b_batch* b, const uint64_t* leaf_indices, uint64_t count,
                    uint64_t* siblings_out /* [count][log2 N - cap_height][4] */);
/* Rows + proofs for FRI query rounds in one gather (fri/prover.rs:187-216 reads them row by row). */
int p2b_batch_open_rows(const p2b_batch* b, const uint64_t* leaf_indices, uint64_t count,
                        uint64_t* rows_out /* [count][leaf_len] */, uint64_t* siblings_out /* or NULL */);

/* ---------------------------------------------------------------------------------------------------
 * Sharded commit (multi-GPU).  The LDE domain splits into 2^rate_bits cosets; in leaf order
 * coset block b is the contiguous leaf range [b*n, (b+1)*n).  A rank commits a contiguous range of blocks from the
 * full coefficient matrix (device, [P][n]): it holds only its own leaf rows, writes its digests into a buffer with
 * the whole tree's layout, and computes digest layers for as long as whole nodes lie inside its leaf range
 * (`top_layer`).  If top_layer reaches the cap (always when 2^cap_height >= number of ranks) the ranks only
 * exchange cap entries; otherwise they exchange their top-layer nodes with export/import and call
 * p2b_batch_finish_layers.  The exchange itself (NCCL all-gather) is done by the caller on the device buffers.
 * ------------------------------------------------------------------------------------------------- */
int p2b_commit_blocks(p2b_ctx* ctx, const uint64_t* d_coeffs, uint32_t n_log, uint64_t P, uint32_t rate_bits,
                      uint32_t cap_height, const uint64_t* d_salt /* device [4][N] or NULL */, uint64_t block_first,
                      uint64_t block_count, p2b_batch** out);
/* Pipelined variant of p2b_commit_blocks for the multi-GPU flow: the coefficient columns arrive in groups (one all-gather
 * round each); every group is low-degree-extended into the leaf rows and absorbed by the leaves' sponges at once, so the
 * exchange of group g+1 overlaps the LDE + hashing of group g.  Groups are absorbed in column order; every group but the
 * last holds a multiple of 8 columns (the sponge rate, hashing.rs:81-104).  After _finish the batch is identical to one
 * produced by p2b_commit_blocks (same leaves, digests up to the local top layer).  No blinding; num_polys > 4. */
int p2b_commit_blocks_begin(p2b_ctx* ctx, uint32_t n_log, uint64_t num_polys, uint32_t rate_bits, uint32_t cap_height,
                            uint64_t block_first, uint64_t block_count, p2b_batch** out);
int p2b_commit_blocks_absorb(p2b_batch* batch, const uint64_t* d_coeff_cols /* [ncols][n] */, uint64_t col0, uint64_t ncols);
int p2b_commit_blocks_finish(p2b_batch* batch);

int p2b_batch_shard_info(const p2b_batch* b, uint64_t* first_leaf, uint64_t* local_leaves, uint32_t* top_layer,
                         uint64_t* top_node_first, uint64_t* top_node_count);
int p2b_batch_export_nodes(const p2b_batch* b, uint32_t layer, uint64_t node_first, uint64_t count,
                           uint64_t* d_out /* device [count][4] */);
int p2b_batch_import_nodes(p2b_batch* b, uint32_t layer, uint64_t node_first, uint64_t count, const uint64_t* d_in);
int p2b_batch_finish_layers(p2b_batch* b, uint32_t from_layer);

/* ---------------------------------------------------------------------------------------------------
 * Single-process multi-device commit (SURVEY.md sections 5 / 8b: "one process, 8 devices").  The reference's caller is
 * ONE prover process (plonky2/src/fri/oracle.rs:279-545 from_values_with_gpu, plonk/prover.rs:239-700 my_prove), so this
 * is the entry a Rust prover binds to use every GPU of a node: `devices` are CUDA ordinals (NULL = 0..n-1), n a power of
 * two <= 2^rate_bits.  The value columns are dealt over the devices in growing exchange rounds, inverse-transformed
 * where they land, pushed to every peer over NVLink (cudaMemcpyPeerAsync, ordered by cross-device events -- there is no
 * host synchronisation inside a commit) and absorbed round by round by the leaves of the coset blocks each device owns;
 * the cap entries are exchanged at the end.  Results are bit-identical to p2b_commit_from_values.  INTEGRATION.md shows
 * the Rust call.
 * ------------------------------------------------------------------------------------------------- */
typedef struct p2b_mgpu p2b_mgpu;
typedef struct p2b_mgpu_batch p2b_mgpu_batch;
int p2b_mgpu_create(const int* devices, int n_dev, p2b_mgpu** out);
void p2b_mgpu_destroy(p2b_mgpu* g);
int p2b_mgpu_device_count(const p2b_mgpu* g);
p2b_ctx* p2b_mgpu_ctx(p2b_mgpu* g, int index);   /* the per-device context (quotient / FRI calls on that device's shard) */
int p2b_mgpu_peer_access(const p2b_mgpu* g);     /* 1 if every pair of devices has direct peer access */
int p2b_mgpu_synchronize(p2b_mgpu* g);
int p2b_mgpu_timer_start(p2b_mgpu* g);           /* CUDA events on every device's stream ... */
int p2b_mgpu_timer_stop_ms(p2b_mgpu* g, float* ms_max); /* ... longest span over the devices */
/* PolynomialBatch::from_values (fri/oracle.rs:709-731): values_host [P][n] column-major (pinned memory overlaps the upload with
 * the transforms; it must stay valid until p2b_mgpu_synchronize); coeffs_host_out NULL or [P][n] (valid after the same).
 * The call enqueues the devices' work from one short-lived host thread per device (joined before it returns; environment
 * variable P2B_MGPU_SINGLE_THREAD=1 keeps it on the caller's thread). */
int p2b_mgpu_commit_from_values(p2b_mgpu* g, const uint64_t* values_host, uint32_t degree_log, uint64_t num_polys,
                                uint32_t rate_bits, uint32_t cap_height, uint64_t* coeffs_host_out, p2b_mgpu_batch** out);
/* The value matrix [P][n] sits on ONE device of the group (index src_index; e.g. Z / partial products computed there): its
 * columns are pushed to their owners over NVLink, then the commit proceeds as above. */
int p2b_mgpu_commit_from_device_values(p2b_mgpu* g, int src_index, const uint64_t* d_values, uint32_t degree_log, uint64_t num_polys,
                                       uint32_t rate_bits, uint32_t cap_height, p2b_mgpu_batch** out);
/* Inputs already resident on the devices.  The columns are dealt in exchange rounds of 8, 8, 16, 32, .. 8*n_dev consecutive
 * columns (p2b_mgpu_round); device `index` holds, for round j, its `per` columns [col0 + index*per, ...) at rows
 * [row0, row0 + per) of its local buffer [sum of per][n] (zero rows where a round is ragged).  The commit transforms them in place. */
int p2b_mgpu_resident_cols(p2b_mgpu* g, int index, uint32_t degree_log, uint64_t num_polys, uint64_t** d_cols_out, uint64_t* rounds_out);
int p2b_mgpu_round(const p2b_mgpu* g, uint64_t num_polys, uint64_t j, uint64_t* col0, uint64_t* width, uint64_t* per, uint64_t* row0);
int p2b_mgpu_commit_resident(p2b_mgpu* g, uint32_t degree_log, uint64_t num_polys, uint32_t rate_bits, uint32_t cap_height,
                             uint64_t* coeffs_host_out, p2b_mgpu_batch** out);
void p2b_mgpu_batch_destroy(p2b_mgpu_batch* b);
int p2b_mgpu_batch_get_info(const p2b_mgpu_batch* b, p2b_batch_info* out);
p2b_batch* p2b_mgpu_batch_shard(p2b_mgpu_batch* b, int index);  /* device `index`'s leaves [index*N/n_dev, (index+1)*N/n_dev) */
int p2b_mgpu_batch_get_cap(const p2b_mgpu_batch* b, uint64_t* out /* [2^cap_height][4] */);
int p2b_mgpu_batch_open_rows(const p2b_mgpu_batch* b, const uint64_t* leaf_indices, uint64_t count, uint64_t* rows_out,
                             uint64_t* siblings_out /* or NULL */);
int p2b_mgpu_batch_get_leaves(const p2b_mgpu_batch* b, uint64_t first_leaf, uint64_t count, uint64_t* out);
/* PolynomialBatch::from_coeffs (fri/oracle.rs:911-977) when every device already holds the coefficient matrix [P][n]
 * (d_coeffs[index] = device `index`'s copy), e.g. the quotient chunks produced by p2b_mgpu_quotient_polys. */
int p2b_mgpu_commit_from_device_coeffs(p2b_mgpu* g, const uint64_t* const* d_coeffs, uint32_t degree_log, uint64_t num_polys,
                                       uint32_t rate_bits, uint32_t cap_height, p2b_mgpu_batch** out);

/* ---------------------------------------------------------------------------------------------------
 * Quotient polynomials: compute_quotient_polys (plonky2/src/plonk/prover.rs:790-1034) for a circuit given as data.
 * The reference CUDA kernel hard-codes one circuit (cuda/plonky2_gpu_impl.cuh:600-685, plonky2_gpu.cu:665-689); here
 * the host passes the part of CommonCircuitData the evaluation reads (circuit_data.rs:270-349).
 *   gates[i]   : `common_data.gates[i]` -- type, its selector polynomial index and group range
 *                (SelectorsInfo, gates/selectors.rs:13-23) and the gate's own parameters:
 *                  NOOP -, CONSTANT p0=num_consts, PUBLIC_INPUT -, ARITHMETIC p0=num_ops, BASE_SUM p0=num_limbs p1=B,
 *                  POSEIDON -, RANDOM_ACCESS p0=bits p1=num_copies p2=num_extra_constants, U32_ARITHMETIC p0=num_ops,
 *                  U32_ADD_MANY p0=num_addends p1=num_ops, U32_RANGE_CHECK p0=num_input_limbs, U32_SUBTRACTION p0=num_ops,
 *                  COMPARISON p0=num_bits p1=num_chunks, ARITHMETIC_EXTENSION / MUL_EXTENSION p0=num_ops,
 *                  REDUCING / REDUCING_EXTENSION p0=num_coeffs, EXPONENTIATION p0=num_power_bits, POSEIDON_MDS -,
 *                  HIGH_ / LOW_DEGREE_INTERPOLATION p0=subgroup_bits (1..8)
 *                  (gates/arithmetic_extension.rs:129, multiplication_extension.rs:122, reducing.rs:160,
 *                  reducing_extension.rs:160, exponentiation.rs:266, poseidon_mds.rs:184, high_degree_interpolation.rs:126,
 *                  low_degree_interpolation.rs:356)
 *   rows       : the three commitments' leaf rows as the commit produces them (row L = LDE point reverse_bits(L)):
 *                wires [num_wires..], zs_partial_products [num_challenges * (1 + num_partial_products)..],
 *                constants_sigmas [num_constants + num_routed_wires..]; trailing salt columns are ignored.
 * Outputs (device, either may be NULL): values [num_challenges][n * 2^ceil(log2 qdf)] = the quotient evaluations on the
 * coset (prover.rs:884-1012), coeffs = their coset_ifft (prover.rs:1014-1021), both canonical.
 * ------------------------------------------------------------------------------------------------- */
enum {
  P2B_GATE_NOOP = 0, P2B_GATE_CONSTANT = 1, P2B_GATE_PUBLIC_INPUT = 2, P2B_GATE_ARITHMETIC = 3, P2B_GATE_BASE_SUM = 4,
  P2B_GATE_POSEIDON = 5, P2B_GATE_RANDOM_ACCESS = 6, P2B_GATE_U32_ARITHMETIC = 7, P2B_GATE_U32_ADD_MANY = 8,
  P2B_GATE_U32_RANGE_CHECK = 9, P2B_GATE_U32_SUBTRACTION = 10, P2B_GATE_COMPARISON = 11,
  /* recursion gate set; an extension element = 2 consecutive wires */
  P2B_GATE_ARITHMETIC_EXTENSION = 12, P2B_GATE_MUL_EXTENSION = 13, P2B_GATE_REDUCING = 14, P2B_GATE_REDUCING_EXTENSION = 15,
  P2B_GATE_EXPONENTIATION = 16, P2B_GATE_POSEIDON_MDS = 17, P2B_GATE_HIGH_DEGREE_INTERPOLATION = 18,
  P2B_GATE_LOW_DEGREE_INTERPOLATION = 19
};
typedef struct {
  uint32_t type, selector_index, group_start, group_end; /* group = gate index range [start, end) of its selector */
  uint32_t p0, p1, p2, reserved;
} p2b_gate;
typedef struct {
  uint32_t degree_bits, rate_bits, quotient_degree_factor, num_challenges;
  uint32_t num_wires, num_routed_wires, num_constants, num_selectors;
  uint32_t num_gates, reserved;
  const p2b_gate* gates;  /* host */
  const uint64_t* k_is;   /* host [num_routed_wires]: common_data.k_is */
} p2b_circuit;

int p2b_quotient_polys(p2b_ctx* ctx, const p2b_circuit* circuit, const p2b_batch* wires, const p2b_batch* zs_partial_products,
                       const p2b_batch* constants_sigmas, const uint64_t* public_inputs_hash /* host [4] */,
                       const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas /* host [num_challenges] */,
                       uint64_t* d_values_out, uint64_t* d_coeffs_out);
int p2b_quotient_polys_rows(p2b_ctx* ctx, const p2b_circuit* circuit, const uint64_t* d_wires_rows, uint64_t wires_stride,
                            const uint64_t* d_zs_pp_rows, uint64_t zs_pp_stride, const uint64_t* d_consts_sigmas_rows,
                            uint64_t consts_sigmas_stride, const uint64_t* public_inputs_hash, const uint64_t* betas,
                            const uint64_t* gammas, const uint64_t* alphas, uint64_t* d_values_out, uint64_t* d_coeffs_out);

/* ---------------------------------------------------------------------------------------------------
 * Permutation argument: all_wires_permutation_partial_products (plonky2/src/plonk/prover.rs:702-786)
 * followed by the re-ordering of prover.rs:112-117, i.e. the matrix the reference commits as zs_partial_products.
 *   d_wires_values  column-major [>= num_routed_wires][n]: the witness (MatrixWitness.wire_values, iop/witness.rs)
 *   d_sigma_values  column-major [num_routed_wires][n]: values of the sigma polynomials on H (prover_data.sigmas, transposed)
 *   k_is, betas, gammas: host arrays ([num_routed_wires], [num_challenges], [num_challenges])
 *   d_out           column-major [num_challenges * ceil(num_routed_wires / quotient_degree_factor)][n]:
 *                   Z_c for every challenge, then the partial products of challenge 0, 1, ... (canonical values)
 * A zero denominator (probability ~ n * num_routed / p for honest challenges) returns P2B_ERR_INVALID "Tried to invert zero",
 * where the reference's batch_multiplicative_inverse panics with the same words (field/src/types.rs:130).  The call
 * synchronises the context's stream before returning (the flag is read back; host arrays may be reused at once).
 * ------------------------------------------------------------------------------------------------- */
int p2b_partial_products_and_zs(p2b_ctx* ctx, const uint64_t* d_wires_values, const uint64_t* d_sigma_values,
                                uint32_t degree_bits, uint32_t num_routed_wires, uint32_t quotient_degree_factor,
                                uint32_t num_challenges, const uint64_t* k_is, const uint64_t* betas, const uint64_t* gammas,
                                uint64_t* d_out);

/* ---------------------------------------------------------------------------------------------------
 * FRI opening proof (SURVEY.md section 8(f) rank 1 + 4): PolynomialBatch::prove_openings (plonky2/src/fri/oracle.rs:1046-1110)
 * -> fri_proof (plonky2/src/fri/prover.rs:23-70), and OpeningSet::new's evaluations (plonky2/src/plonk/proof.rs:305-334).
 * The reference runs all of it on the CPU; here the committed batches never leave the device:
 *   alpha <- challenger; final_poly = X * sum_i alpha^(k_i) (F_i - F_i(z_i)) / (X - z_i), F_i = sum_j alpha^j f_ij
 *   (reduce_polys_base + divide_by_linear + shift_poly, util/reducing.rs:87-111, field/src/polynomial/division.rs:75-88);
 *   coset LDE over the quadratic extension; per reduction: Merkle tree over arity-sized bit-reversed chunks, observe_cap,
 *   beta, fold the coefficients, coset FFT on shift^arity (prover.rs:76-120); final polynomial; proof-of-work grinding
 *   (prover.rs:123-171); query indices and the rows + Merkle paths of every tree (prover.rs:173-260).
 * Extension elements are pairs (c0, c1) of F[X]/(X^2 - 7) (field/src/goldilocks_extensions.rs:14-28).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t sponge_state[12]; /* iop/challenger.rs:16-21 */
  uint64_t input_buffer[8];
  uint64_t output_buffer[8]; /* challenges are popped from the END (Vec::pop, challenger.rs:90-92) */
  uint32_t input_len, output_len;
} p2b_challenger;
typedef struct {
  uint32_t oracle_index, polynomial_index; /* FriPolynomialInfo, fri/structure.rs:46-52 */
} p2b_fri_poly_info;
typedef struct {
  uint64_t point[2];                     /* FriBatchInfo.point, fri/structure.rs:34-38 */
  const p2b_fri_poly_info* polynomials;  /* host */
  uint32_t num_polynomials, reserved;
} p2b_fri_batch_info;
typedef struct {
  uint32_t degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds; /* FriParams / FriConfig, fri/mod.rs:17-75 */
  uint32_t num_reductions;
  const uint32_t* reduction_arity_bits;  /* host [num_reductions] */
} p2b_fri_params;
typedef struct {
  uint32_t num_reductions, num_query_rounds, num_oracles, cap_height;
  uint64_t final_poly_len;  /* extension elements */
  uint64_t lde_size;
} p2b_fri_proof_info;
typedef struct p2b_fri_proof p2b_fri_proof;

/* f(point) for every polynomial of a committed batch (eval_commitment, proof.rs:313-319); out: host [num_polys][2]. */
int p2b_eval_openings(p2b_ctx* ctx, const p2b_batch* batch, const uint64_t point[2], uint64_t* out);
/* prove_openings.  `challenger` (host) is the transcript state on entry and is advanced exactly as the reference advances
 * it (alpha, caps/betas, final polynomial, PoW witness + response, query challenges).  Oracles must be whole (unsharded)
 * batches of the same degree with rate_bits == params->rate_bits. */
int p2b_fri_prove_openings(p2b_ctx* ctx, const p2b_batch* const* oracles, uint32_t num_oracles,
                           const p2b_fri_batch_info* batches, uint32_t num_batches, p2b_challenger* challenger,
                           const p2b_fri_params* params, p2b_fri_proof** out);
void p2b_fri_proof_destroy(p2b_fri_proof* proof);
int p2b_fri_proof_get_info(const p2b_fri_proof* proof, p2b_fri_proof_info* out);
/* commit_phase_merkle_caps[round]: [2^cap_height][4] */
int p2b_fri_proof_get_cap(const p2b_fri_proof* proof, uint32_t round, uint64_t* out);
int p2b_fri_proof_get_final_poly(const p2b_fri_proof* proof, uint64_t* out /* [final_poly_len][2] */);
int p2b_fri_proof_get_pow_witness(const p2b_fri_proof* proof, uint64_t* out);
int p2b_fri_proof_get_query_indices(const p2b_fri_proof* proof, uint64_t* out /* [num_query_rounds] */);
/* FriInitialTreeProof of every query for one oracle: rows [Q][leaf_len] (salt columns included, as MerkleTree::get returns
 * them), siblings [Q][lde_bits - cap_height][4]. */
int p2b_fri_proof_get_initial(const p2b_fri_proof* proof, uint32_t oracle, uint64_t* rows_out, uint64_t* siblings_out);
/* FriQueryStep of every query for one reduction: evals [Q][arity][2], siblings [Q][depth][4];
 * depth = log2(tree leaves) - cap_height is returned in *depth_out when non-NULL (either output may be NULL). */
int p2b_fri_proof_get_step(const p2b_fri_proof* proof, uint32_t round, uint64_t* evals_out, uint64_t* siblings_out,
                           uint32_t* depth_out);
/* Transcript values, for parity tests: what = 0 alpha [2], 1 betas [num_reductions][2], 2 the polynomial that enters FRI
 * (oracle.rs:1084) [n][2], 3 PoW response [1]. */
int p2b_fri_proof_get_debug(const p2b_fri_proof* proof, uint32_t what, uint64_t* out);

/* ---------------------------------------------------------------------------------------------------
 * Building blocks (device pointers unless stated).  Each mirrors one reference function.
 * ------------------------------------------------------------------------------------------------- */
/* values.into_par_iter().map(|v| v.ifft())  (fri/oracle.rs:717-721, field/src/fft.rs:73-103).
 * d_src / d_dst column-major [P][n]; may alias. */
int p2b_ifft_batch(p2b_ctx* ctx, const uint64_t* d_src, uint64_t* d_dst, uint32_t n_log, uint64_t P);
/* lde_values + transpose + reverse_index_bits (fri/oracle.rs:979-1004, 942-952): d_coeffs [P][n] ->
 * d_leaves rows [N][row_stride] (columns col0 .. col0+P of each row). */
int p2b_lde_leaves(p2b_ctx* ctx, const uint64_t* d_coeffs, uint32_t n_log, uint64_t P, uint32_t rate_bits,
                   uint64_t* d_leaves, uint64_t row_stride, uint64_t col0);
/* MerkleTree::new on leaf rows (merkle_tree.rs:283-319).  Element (row, col) of the leaves is at
 * d_leaves[row * row_stride + col * col_stride]. */
int p2b_merkle_tree(p2b_ctx* ctx, const uint64_t* d_leaves, uint64_t num_leaves, uint64_t leaf_len,
                    uint64_t row_stride, uint64_t col_stride, uint32_t cap_height, uint64_t* d_digests,
                    uint64_t* d_cap);
/* F::poseidon on `count` independent 12-word states, in place (hash/poseidon.rs:590-606). */
int p2b_poseidon_permute(p2b_ctx* ctx, uint64_t* d_states, uint64_t count);
/* Element-wise field ops for the parity tests: op 0 add, 1 sub, 2 mul; out[i] = a[i] op b[i] (canonical). */
int p2b_field_op(p2b_ctx* ctx, int op, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint64_t count);
/* Deterministic synthetic input: out[i] = splitmix64(seed, first_index + i) rejected >= p. */
int p2b_fill_synthetic(p2b_ctx* ctx, uint64_t* d_out, uint64_t count, uint64_t seed, uint64_t first_index);

/* Device memory helpers so that non-CUDA hosts (ctypes, Rust FFI) need no second allocator.  p2b_malloc / p2b_free take
 * from / return to the device's stream-ordered pool, ordered on the context's stream (a prove() stage's buffers cost
 * microseconds after the first proof; nothing is handed back to the driver until the process ends): use the pointer in
 * calls on the same context, or synchronise the context before handing it to other streams. */
int p2b_malloc(p2b_ctx* ctx, uint64_t bytes, void** out);
/* How many pool allocations of this process were repeated after a transient "out of memory" from the stream-ordered
 * allocator (diagnostics; 0 in a healthy single-device run). */
unsigned long long p2b_debug_pool_retries(void);
int p2b_free(p2b_ctx* ctx, void* ptr);
int p2b_malloc_host(uint64_t bytes, void** out); /* pinned */
int p2b_free_host(void* ptr);
int p2b_memcpy_h2d(p2b_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes);
int p2b_memcpy_d2h(p2b_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes);
/* Time on the context's stream: event pair helpers returning milliseconds (bench.py uses these so the
 * timing is taken on the stream the kernels are launched on). */
int p2b_timer_start(p2b_ctx* ctx);
int p2b_timer_stop_ms(p2b_ctx* ctx, float* ms_out);

/* ---------------------------------------------------------------------------------------------------
 * Multi-device quotient polynomials and FRI opening proofs (the stages of prove() after a p2b_mgpu commit)
 * ------------------------------------------------------------------------------------------------- */
/* compute_quotient_polys (plonk/prover.rs:790-1034) over sharded batches: every device evaluates the points whose leaf rows
 * it owns (no row crosses NVLink; needs n_dev <= 2^quotient_degree_bits), the compact value vectors are exchanged, and every
 * device ends with the coefficients [num_challenges][lde_size] in d_coeffs_out[index] (device memory of that device,
 * caller-allocated) -- ready for p2b_mgpu_commit_from_device_coeffs as [num_challenges * qdf][n] chunks. */
int p2b_mgpu_quotient_polys(p2b_mgpu* g, const p2b_circuit* circuit, const p2b_mgpu_batch* wires, const p2b_mgpu_batch* zs_partial_products,
                            const p2b_mgpu_batch* constants_sigmas, const uint64_t* public_inputs_hash, const uint64_t* betas,
                            const uint64_t* gammas, const uint64_t* alphas, uint64_t* const* d_coeffs_out);
/* OpeningSet::new's eval_commitment and PolynomialBatch::prove_openings (plonk/proof.rs:313-319, fri/oracle.rs:1046-1110) over
 * sharded oracles: coefficient work on the first device (every shard holds a complete coefficient copy), query rows and
 * Merkle paths from the devices that own the leaves.  The proof object is read with the p2b_fri_proof_* getters. */
int p2b_mgpu_eval_openings(p2b_mgpu* g, const p2b_mgpu_batch* b, const uint64_t point[2], uint64_t* out);
int p2b_mgpu_fri_prove_openings(p2b_mgpu* g, const p2b_mgpu_batch* const* oracles, uint32_t num_oracles,
                                const p2b_fri_batch_info* batches, uint32_t num_batches, p2b_challenger* challenger,
                                const p2b_fri_params* params, p2b_fri_proof** out);

/* ---------------------------------------------------------------------------------------------------
 * Reference-compatible symbols (cuda/src/lib.rs:52-145).  Same names, argument order, in-place device
 * layout and error convention (struct returned by value: CUDA error code + strdup'd message or NULL, which
 * the Rust side frees, lib.rs:20-51).  `ctx` points at {cudaStream_t stream; cudaStream_t stream2;}
 * (cuda/plonky2_gpu.cu:4-7).  root tables / shift powers / n_inv arguments are accepted and ignored: the
 * library owns its twiddle tables.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int code;
  char* message;
} p2b_rust_error; /* == cuda::Error (lib.rs:20-25) == RustError (plonky2_gpu.cu:19-31) */

typedef struct {
  const void* ptr;
  int len;
} p2b_data_slice; /* == DataSlice (lib.rs:52-56) */

#ifndef P2B_NO_COMPAT_SYMBOLS
void init(void); /* lib.rs:59 (declared there, never defined by the reference) */
p2b_rust_error ifft(uint64_t* d_values_flatten, int poly_num, int values_num_per_poly, int log_len,
                    const uint64_t* d_root_table, const uint64_t* p_inv, void* ctx); /* lib.rs:61-69 */
p2b_rust_error build_merkle_tree(uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly,
                                 int log_len, int rate_bits, int salt_size, int cap_height,
                                 int pad_extvalues_len, void* ctx); /* lib.rs:71-81 */
p2b_rust_error merkle_tree_from_values(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num,
                                       int values_num_per_poly, int log_len, const uint64_t* d_root_table,
                                       const uint64_t* d_root_table2, const uint64_t* d_shift_powers,
                                       const uint64_t* p_inv, int rate_bits, int salt_size, int cap_height,
                                       int pad_extvalues_len, void* ctx); /* lib.rs:83-98 */
p2b_rust_error merkle_tree_from_coeffs(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num,
                                       int values_num_per_poly, int log_len, const uint64_t* d_root_table,
                                       const uint64_t* d_root_table2, const uint64_t* d_shift_powers,
                                       int rate_bits, int salt_size, int cap_height, int pad_extvalues_len,
                                       void* ctx); /* lib.rs:100-114 */
p2b_rust_error transpose(uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly, int rate_bits,
                         int salt_size, int pad_extvalues_len, void* ctx); /* plonky2_gpu.cu:192-215 */
/* plonky2_gpu.cu:88-136 (defined by the reference, not declared in lib.rs): coset LDE of the coefficient columns into the
 * work area as a column-major [poly_num][N] matrix in natural point order. */
p2b_rust_error fft_blinding(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly,
                            int log_len, const uint64_t* d_root_table2, const uint64_t* d_shift_powers, int rate_bits,
                            int pad_extvalues_len, void* ctx);
/* lib.rs:117-143.  The reference kernel is compiled for ONE circuit; this one evaluates the circuit registered with
 * p2b_compat_set_circuit().  d_outs: [N][num_challenges] values, d_quotient_polys: [num_challenges][N] coefficients. */
p2b_rust_error compute_quotient_polys(const uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly, int log_len,
                                      const uint64_t* d_root_table2, const uint64_t* d_shift_inv_powers, int rate_bits,
                                      int salt_size, const p2b_data_slice* zs_partial_products_commitment_leaves,
                                      const p2b_data_slice* constants_sigmas_commitment_leaves, void* d_outs,
                                      void* d_quotient_polys, const p2b_data_slice* points,
                                      const p2b_data_slice* z_h_on_coset_evals, const p2b_data_slice* z_h_on_coset_inverses,
                                      const p2b_data_slice* k_is, const p2b_data_slice* alphas, const p2b_data_slice* betas,
                                      const p2b_data_slice* gammas, void* ctx);
#endif
/* Registers the circuit (and the public-inputs hash) the legacy compute_quotient_polys symbol evaluates. */
int p2b_compat_set_circuit(const p2b_circuit* circuit, const uint64_t* public_inputs_hash /* host [4] */);

#ifdef __cplusplus
}
#endif
#endif /* PLONKY2_B200_H */
