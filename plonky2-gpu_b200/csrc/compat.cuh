// compat.cuh -- the reference's own C symbols (cuda/src/lib.rs:52-145, cuda/plonky2_gpu.cu:57-785) on top of
// the B200 kernels, so that the reference's Rust crate `plonky2_cuda` links against this library unchanged.
// Included at the end of plonky2_b200.cu (single translation unit).
//
// Device-memory contract kept from the reference (plonky2/src/fri/oracle.rs:316-455,
// cuda/plonky2_gpu.cu:435-606), element = u64, base = the pointer the caller passes:
//   in : coefficients (or values for ifft) column-major at base[0 .. n*P)
//   out: leaf-major LDE  base[0 .. N*P)            (row L = reference leaf L, bit-reversed order)
//        work area       base[pad .. 2*pad)        (pad = pad_extvalues_len = N*(P+salt))
//        digests, cap    base[2*pad .. 2*pad + 4*(num_digests + 2^cap_height))
// `ctx` is the reference's {cudaStream_t stream, stream2} pair (plonky2_gpu.cu:4-7); stream2 carries the
// caller's D2H copy of the coefficients, which must finish before base[0 .. n*P) is overwritten
// (plonky2_gpu.cu:586).  Unlike the reference the shims do not printf, and they report launch errors.
#pragma once

struct RefStreams {
  cudaStream_t stream;
  cudaStream_t stream2;
};

static std::mutex g_compat_mu;
static p2b_ctx* g_compat_ctx[64];

static p2b_rust_error rust_ok() { return p2b_rust_error{0, nullptr}; }
static p2b_rust_error rust_err(int rc) {
  // code: a cudaError_t-like non-zero value; message strdup'd, freed by the Rust side (lib.rs:27-36)
  const char* m = p2b_last_error();
  return p2b_rust_error{rc == P2B_ERR_OOM ? (int)cudaErrorMemoryAllocation : (rc == P2B_ERR_INVALID ? (int)cudaErrorInvalidValue : (int)cudaErrorUnknown),
                        m && *m ? strdup(m) : nullptr};
}

// One lazily created context per device for the legacy entry points.  A legacy call holds g_compat_mu from entry to
// return (CompatCall): the shared context's main stream is replaced by the caller's stream for the duration of the call
// (the reference launches everything on ctx->stream), and its scratch buffer and events are not reentrant, so two host
// threads calling the legacy symbols are serialised instead of corrupting each other (the reference's own caller is
// single-threaded per CudaInvContext, plonky2/src/fri/oracle.rs:279-545).
struct CompatCall {
  std::unique_lock<std::mutex> lk;
  p2b_ctx* c = nullptr;
  cudaStream_t saved = nullptr;
  int rc = P2B_OK;
  explicit CompatCall(void* ref_ctx) : lk(g_compat_mu) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { rc = fail(P2B_ERR_CUDA, "cudaGetDevice failed"); return; }
    if (dev < 0 || dev >= 64) { rc = fail(P2B_ERR_INVALID, "device index %d", dev); return; }
    if (!g_compat_ctx[dev]) {
      rc = p2b_ctx_create(dev, &g_compat_ctx[dev]);
      if (rc != P2B_OK) return;
    }
    c = g_compat_ctx[dev];
    saved = c->stream;
    if (ref_ctx) {
      cudaStream_t s = static_cast<RefStreams*>(ref_ctx)->stream;
      if (s) c->stream = s;
    }
  }
  ~CompatCall() {
    if (c) c->stream = saved;
  }
  // wait for the caller's stream (the reference synchronises after every kernel) and translate the status
  p2b_rust_error finish(int status) {
    if (status == P2B_OK && c && cudaStreamSynchronize(c->stream) != cudaSuccess) status = fail(P2B_ERR_CUDA, "stream synchronize failed");
    return status == P2B_OK ? rust_ok() : rust_err(status);
  }
};

extern "C" void init(void) { CompatCall call(nullptr); }

// ifft: in-place inverse NTT of poly_num columns (cuda/plonky2_gpu.cu:70-86; oracle.rs:394-401)
extern "C" p2b_rust_error ifft(uint64_t* d_values_flatten, int poly_num, int values_num_per_poly, int log_len,
                               const uint64_t* d_root_table, const uint64_t* p_inv, void* ctx) {
  (void)d_root_table;
  (void)p_inv;
  if (!d_values_flatten || poly_num <= 0 || log_len < 0 || values_num_per_poly != (1 << log_len)) {
    fail(P2B_ERR_INVALID, "ifft: bad arguments");
    return rust_err(P2B_ERR_INVALID);
  }
  CompatCall call(ctx);
  if (call.rc != P2B_OK) return rust_err(call.rc);
  return call.finish(p2b_ifft_batch(call.c, d_values_flatten, d_values_flatten, (u32)log_len, (u64)poly_num));
}

static int compat_from_coeffs(p2b_ctx* c, RefStreams* rs, u64* base, int poly_num, int log_len, int rate_bits,
                              int salt_size, int cap_height, long long pad) {
  // Blinding (salt_size = SALT_SIZE = 4, oracle.rs:302): the reference never uploads salt values -- its hash_leaves /
  // transpose kernels read columns [poly_num, poly_num + salt_size) of the column-major work area as they find them
  // (plonky2_gpu.cu:552-600: those columns are neither written by the LDE nor bit-reversed), i.e. leaf L gets
  // work[(poly_num + s) * N + L].  The same words are used here (canonicalised), so a caller that fills them gets
  // exactly its salt and one that does not gets the same "whatever the device buffer held" the reference hashes.
  if (salt_size < 0 || salt_size > 64) return fail(P2B_ERR_INVALID, "salt_size %d", salt_size);
  const u64 P = (u64)poly_num, n = (u64)1 << log_len, N = n << rate_bits;
  if ((u64)pad < N * (P + (u64)salt_size)) return fail(P2B_ERR_INVALID, "pad_extvalues_len smaller than the LDE matrix");
  const u64 ncap = (u64)1 << cap_height;
  u64* work = base + pad;
  u64* digests = work + N * (P + (u64)salt_size);  // d_digest_buf, plonky2_gpu.cu:552
  u64* cap = digests + 4 * 2 * (N - ncap);
  const u64* coeffs = base;
  if (make_plan((u32)log_len).n_strided == 0) {
    // tiny transforms have a single pass that would read the coefficients while other CTAs overwrite them
    CUDA_TRY(cudaMemcpyAsync(work, base, n * P * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
    coeffs = work;
  }
  // intermediates live in the work area; coset blocks are produced in descending order so that block 0,
  // whose rows overwrite the coefficients, is written last and only after stream2's copy has finished.
  return lde_and_merkle(c, coeffs, (u32)log_len, P, (u32)rate_bits, (u32)cap_height, nullptr, work, base, P + (u64)salt_size, digests, cap,
                        true, rs ? rs->stream2 : nullptr, 0, (u64)1 << rate_bits, nullptr, salt_size ? work + P * N : nullptr,
                        (u32)salt_size);
}

// merkle_tree_from_coeffs (cuda/plonky2_gpu.cu:435-606; oracle.rs:409-422, 599-627)
extern "C" p2b_rust_error merkle_tree_from_coeffs(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num,
                                                  int values_num_per_poly, int log_len, const uint64_t* d_root_table,
                                                  const uint64_t* d_root_table2, const uint64_t* d_shift_powers,
                                                  int rate_bits, int salt_size, int cap_height, int pad_extvalues_len,
                                                  void* ctx) {
  (void)d_root_table;
  (void)d_root_table2;
  (void)d_shift_powers;
  if (!d_values_flatten || d_ext_values_flatten != d_values_flatten || poly_num <= 0 || log_len < 0 ||
      values_num_per_poly != (1 << log_len) || rate_bits < 0 || cap_height < 0 || cap_height > log_len + rate_bits) {
    fail(P2B_ERR_INVALID, "merkle_tree_from_coeffs: bad arguments (the reference passes the same pointer for values and ext_values)");
    return rust_err(P2B_ERR_INVALID);
  }
  CompatCall call(ctx);
  if (call.rc != P2B_OK) return rust_err(call.rc);
  return call.finish(compat_from_coeffs(call.c, static_cast<RefStreams*>(ctx), d_values_flatten, poly_num, log_len, rate_bits, salt_size,
                                        cap_height, pad_extvalues_len));
}

// merkle_tree_from_values (declared lib.rs:83-98; the reference's body is `assert(0)`, plonky2_gpu.cu:228):
// implemented here as ifft followed by merkle_tree_from_coeffs, which is what its dead code intended.
extern "C" p2b_rust_error merkle_tree_from_values(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num,
                                                  int values_num_per_poly, int log_len, const uint64_t* d_root_table,
                                                  const uint64_t* d_root_table2, const uint64_t* d_shift_powers,
                                                  const uint64_t* p_inv, int rate_bits, int salt_size, int cap_height,
                                                  int pad_extvalues_len, void* ctx) {
  p2b_rust_error e = ifft(d_values_flatten, poly_num, values_num_per_poly, log_len, d_root_table, p_inv, ctx);
  if (e.code != 0) return e;
  return merkle_tree_from_coeffs(d_values_flatten, d_ext_values_flatten, poly_num, values_num_per_poly, log_len, d_root_table,
                                 d_root_table2, d_shift_powers, rate_bits, salt_size, cap_height, pad_extvalues_len, ctx);
}

// build_merkle_tree (lib.rs:71-81; plonky2_gpu.cu:138-189): the reference takes the column-major, natural-order
// LDE in the work area base[pad ..], bit-reverses it in place, and builds digests + cap behind it.  Here the
// leaves are hashed straight from the column-major matrix with the bit-reversed row index, then the matrix is
// permuted so the work area ends in the state the reference leaves it in.
__global__ void bitrev_rows_colmajor_kernel(u64* __restrict__ m, u64 N, u32 log_N, u64 ncols) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 col = blockIdx.y;
  if (i >= N) return;
  u64 j = log_N ? (__brevll(i) >> (64 - log_N)) : 0;
  if (i < j) {
    u64* p = m + col * N;
    u64 a = p[i], b = p[j];
    p[i] = b;
    p[j] = a;
  }
}
extern "C" p2b_rust_error build_merkle_tree(uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly, int log_len,
                                            int rate_bits, int salt_size, int cap_height, int pad_extvalues_len, void* ctx) {
  if (!d_ext_values_flatten || poly_num <= 0 || log_len < 0 || values_num_per_poly != (1 << log_len) || rate_bits < 0 ||
      salt_size < 0 || cap_height < 0 || cap_height > log_len + rate_bits) {
    fail(P2B_ERR_INVALID, "build_merkle_tree: bad arguments");
    return rust_err(P2B_ERR_INVALID);
  }
  CompatCall call(ctx);
  if (call.rc != P2B_OK) return rust_err(call.rc);
  p2b_ctx* c = call.c;
  int rc = P2B_OK;
  const u64 N = (u64)values_num_per_poly << rate_bits, cols = (u64)poly_num + salt_size;
  const u32 log_N = (u32)(log_len + rate_bits);
  u64* m = d_ext_values_flatten + pad_extvalues_len;
  u64* digests = m + N * cols;
  u64* cap = digests + 4 * 2 * (N - ((u64)1 << cap_height));
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)cols);
  bitrev_rows_colmajor_kernel<<<grid, 256, 0, c->stream>>>(m, N, log_N, cols);
  c->launches++;
  if (cudaGetLastError() != cudaSuccess) rc = fail(P2B_ERR_CUDA, "bit-reversal launch failed");
  if (rc == P2B_OK) rc = p2b_merkle_tree(c, m, N, cols, 1, N, (u32)cap_height, digests, cap);
  return call.finish(rc);
}

// transpose (plonky2_gpu.cu:192-215): column-major [cols][N] in the work area -> row-major [N][cols] at base.
__global__ void transpose_to_rows_kernel(const u64* __restrict__ src, u64* __restrict__ dst, u64 N, u64 cols) {
  __shared__ u64 tile[32][33];
  u64 r0 = (u64)blockIdx.x * 32, c0 = (u64)blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    u64 c = c0 + k, r = r0 + threadIdx.x;
    if (c < cols && r < N) tile[k][threadIdx.x] = src[c * N + r];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    u64 r = r0 + k, c = c0 + threadIdx.x;
    if (c < cols && r < N) dst[r * cols + c] = tile[threadIdx.x][k];
  }
}
extern "C" p2b_rust_error transpose(uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly, int rate_bits,
                                    int salt_size, int pad_extvalues_len, void* ctx) {
  if (!d_ext_values_flatten || poly_num <= 0 || values_num_per_poly <= 0 || rate_bits < 0 || salt_size < 0) {
    fail(P2B_ERR_INVALID, "transpose: bad arguments");
    return rust_err(P2B_ERR_INVALID);
  }
  CompatCall call(ctx);
  if (call.rc != P2B_OK) return rust_err(call.rc);
  p2b_ctx* c = call.c;
  int rc = P2B_OK;
  const u64 N = (u64)values_num_per_poly << rate_bits, cols = (u64)poly_num + salt_size;
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
  transpose_to_rows_kernel<<<grid, block, 0, c->stream>>>(d_ext_values_flatten + pad_extvalues_len, d_ext_values_flatten, N, cols);
  c->launches++;
  if (cudaGetLastError() != cudaSuccess) rc = fail(P2B_ERR_CUDA, "transpose launch failed");
  return call.finish(rc);
}

// fft_blinding (plonky2_gpu.cu:88-136; defined by the reference but not declared in lib.rs): coset LDE of the coefficient
// columns at base[0 .. n*P) into the work area base[pad ..] as a column-major [P][N] matrix in NATURAL point order (no
// bit reversal, no transpose, no hashing -- the caller follows it with build_merkle_tree + transpose).
__global__ void leaves_to_colmajor_natural_kernel(const u64* __restrict__ leaves, u64* __restrict__ dst, u64 N, u32 log_N, u64 cols) {
  __shared__ u64 tile[32][33];
  u64 r0 = (u64)blockIdx.x * 32, c0 = (u64)blockIdx.y * 32;  // r = natural point index
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    u64 r = r0 + k, c = c0 + threadIdx.x;
    if (c < cols && r < N) tile[k][threadIdx.x] = leaves[(log_N ? (__brevll(r) >> (64 - log_N)) : 0) * cols + c];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    u64 c = c0 + k, r = r0 + threadIdx.x;
    if (c < cols && r < N) dst[c * N + r] = tile[threadIdx.x][k];
  }
}
extern "C" p2b_rust_error fft_blinding(uint64_t* d_values_flatten, uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly,
                                       int log_len, const uint64_t* d_root_table2, const uint64_t* d_shift_powers, int rate_bits,
                                       int pad_extvalues_len, void* ctx) {
  (void)d_root_table2;
  (void)d_shift_powers;
  if (!d_values_flatten || !d_ext_values_flatten || poly_num <= 0 || log_len < 0 || values_num_per_poly != (1 << log_len) || rate_bits < 0) {
    fail(P2B_ERR_INVALID, "fft_blinding: bad arguments");
    return rust_err(P2B_ERR_INVALID);
  }
  CompatCall call(ctx);
  if (call.rc != P2B_OK) return rust_err(call.rc);
  p2b_ctx* c = call.c;
  const u64 P = (u64)poly_num, N = (u64)values_num_per_poly << rate_bits;
  const u32 log_N = (u32)(log_len + rate_bits);
  u64* rows = nullptr;
  int rc = P2B_OK;
  if (pool_alloc(&rows, N * P * sizeof(u64), c->stream) != cudaSuccess) rc = fail(P2B_ERR_OOM, "fft_blinding: scratch allocation failed");
  if (rc == P2B_OK) rc = p2b_lde_leaves(c, d_values_flatten, (u32)log_len, P, (u32)rate_bits, rows, P, 0);
  if (rc == P2B_OK) {
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((P + 31) / 32)), block(32, 8);
    leaves_to_colmajor_natural_kernel<<<grid, block, 0, c->stream>>>(rows, d_ext_values_flatten + pad_extvalues_len, N, log_N, P);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = fail(P2B_ERR_CUDA, "fft_blinding: launch failed");
  }
  if (rows) cudaFreeAsync(rows, c->stream);
  return call.finish(rc);
}

// compute_quotient_polys (lib.rs:117-143; plonky2_gpu.cu:609-783).  The reference's kernel is compiled for one circuit
// (literal gate list, constraint count and public-input hash, plonky2_gpu.cu:665-689, plonky2_gpu_impl.cuh:600-685), so its
// C signature carries no circuit description.  Here the host registers the circuit once with p2b_compat_set_circuit()
// (the data CommonCircuitData already holds) and the legacy symbol evaluates THAT circuit:
//   d_ext_values_flatten : wires leaves (leaf-major, leaf_len = poly_num + salt_size)
//   zs_partial_products_commitment_leaves / constants_sigmas_commitment_leaves : device rows, len = N * leaf_len
//   d_outs : [N][num_challenges] quotient values (the reference's point-major layout, plonky2_gpu_impl.cuh:874-875)
//   d_quotient_polys : [num_challenges][N] coefficients after iNTT and the g^-i scaling (plonky2_gpu.cu:735-760)
//   alphas / betas / gammas / k_is : DEVICE slices, as the Rust side uploads them (prover.rs:429-533)
//   points / z_h_on_coset_* / root_table2 / shift_inv_powers : accepted and ignored (recomputed by the library)
struct CompatCircuit {
  bool set = false;
  p2b_circuit circ{};
  std::vector<p2b_gate> gates;
  std::vector<u64> k_is;
  u64 pih[4] = {0, 0, 0, 0};
};
static CompatCircuit g_compat_circuit;

extern "C" int p2b_compat_set_circuit(const p2b_circuit* circuit, const uint64_t* public_inputs_hash) {
  if (!circuit || !public_inputs_hash) return fail(P2B_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lk(g_compat_mu);
  CompatCircuit& cc = g_compat_circuit;
  cc.circ = *circuit;
  cc.gates.assign(circuit->gates, circuit->gates + circuit->num_gates);
  cc.k_is.assign(circuit->k_is, circuit->k_is + circuit->num_routed_wires);
  cc.circ.gates = cc.gates.data();
  cc.circ.k_is = cc.k_is.data();
  for (int i = 0; i < 4; i++) cc.pih[i] = public_inputs_hash[i];
  cc.set = true;
  return P2B_OK;
}

extern "C" p2b_rust_error compute_quotient_polys(const uint64_t* d_ext_values_flatten, int poly_num, int values_num_per_poly,
                                                 int log_len, const uint64_t* d_root_table2, const uint64_t* d_shift_inv_powers,
                                                 int rate_bits, int salt_size, const p2b_data_slice* zs_pp_leaves,
                                                 const p2b_data_slice* consts_sigmas_leaves, void* d_outs, void* d_quotient_polys,
                                                 const p2b_data_slice* points, const p2b_data_slice* z_h_on_coset_evals,
                                                 const p2b_data_slice* z_h_on_coset_inverses, const p2b_data_slice* k_is,
                                                 const p2b_data_slice* alphas, const p2b_data_slice* betas,
                                                 const p2b_data_slice* gammas, void* ctx) {
  (void)d_root_table2; (void)d_shift_inv_powers; (void)points; (void)z_h_on_coset_evals; (void)z_h_on_coset_inverses; (void)k_is;
  CompatCall call(ctx);   // also guards g_compat_circuit against a concurrent p2b_compat_set_circuit
  if (call.rc != P2B_OK) return rust_err(call.rc);
  CompatCircuit& cc = g_compat_circuit;
  if (!cc.set) {
    fail(P2B_ERR_INVALID, "compute_quotient_polys: no circuit registered; call p2b_compat_set_circuit() first (the reference kernel "
                          "hard-codes its circuit, this library takes it as data)");
    return rust_err(P2B_ERR_INVALID);
  }
  const p2b_circuit& circ = cc.circ;
  if (!d_ext_values_flatten || !zs_pp_leaves || !consts_sigmas_leaves || !d_quotient_polys || !alphas || !betas || !gammas ||
      log_len != (int)circ.degree_bits || rate_bits != (int)circ.rate_bits || values_num_per_poly != (1 << log_len) ||
      poly_num < (int)circ.num_wires || alphas->len != (int)circ.num_challenges || betas->len != (int)circ.num_challenges ||
      gammas->len != (int)circ.num_challenges) {
    fail(P2B_ERR_INVALID, "compute_quotient_polys: arguments do not match the registered circuit");
    return rust_err(P2B_ERR_INVALID);
  }
  const u64 N = (u64)values_num_per_poly << rate_bits;
  if (zs_pp_leaves->len <= 0 || consts_sigmas_leaves->len <= 0 || (u64)zs_pp_leaves->len % N || (u64)consts_sigmas_leaves->len % N) {
    fail(P2B_ERR_INVALID, "compute_quotient_polys: leaf slices are not a multiple of the LDE size");
    return rust_err(P2B_ERR_INVALID);
  }
  p2b_ctx* c = call.c;
  int rc = P2B_OK;
  u64 h[3][quotient::MAX_CHALLENGES];
  const p2b_data_slice* sl[3] = {alphas, betas, gammas};
  for (int k = 0; k < 3 && rc == P2B_OK; k++)
    if (cudaMemcpyAsync(h[k], sl[k]->ptr, circ.num_challenges * sizeof(u64), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = fail(P2B_ERR_CUDA, "copy of challenges failed");
  if (rc == P2B_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(P2B_ERR_CUDA, "stream synchronize failed");
  if (rc == P2B_OK)
    rc = quotient_impl(c, &circ, d_ext_values_flatten, (u64)poly_num + salt_size, (const u64*)zs_pp_leaves->ptr,
                       (u64)zs_pp_leaves->len / N, (const u64*)consts_sigmas_leaves->ptr, (u64)consts_sigmas_leaves->len / N, cc.pih,
                       h[1], h[2], h[0], nullptr, (u64*)d_quotient_polys, (u64*)d_outs);
  return call.finish(rc);
}
