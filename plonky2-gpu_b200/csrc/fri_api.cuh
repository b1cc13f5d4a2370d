// FRI opening proof: host orchestration + C ABI (included by plonky2_b200.cu; kernels in fri.cuh).
// Reference: PolynomialBatch::prove_openings (plonky2/src/fri/oracle.rs:1046-1110), fri_proof and its phases
// (plonky2/src/fri/prover.rs:23-260), OpeningSet::new (plonky2/src/plonk/proof.rs:305-334).
#pragma once

namespace hostf {
struct E2h {
  u64 a, b;
};
static inline u64 addm(u64 a, u64 b) { return (u64)(((u128)a + b) % gl::P); }
static inline E2h emul(E2h x, E2h y) {
  return E2h{addm(mul(x.a, y.a), mul(7, mul(x.b, y.b))), addm(mul(x.a, y.b), mul(x.b, y.a))};
}
static inline E2h epow(E2h x, u64 e) {
  E2h r{1, 0};
  while (e) {
    if (e & 1) r = emul(r, x);
    x = emul(x, x);
    e >>= 1;
  }
  return r;
}
}  // namespace hostf

static_assert(sizeof(fri::Challenger) == sizeof(p2b_challenger), "device challenger must mirror p2b_challenger");

struct p2b_fri_proof {
  p2b_fri_proof_info info{};
  std::vector<u32> arity_bits;
  std::vector<std::vector<u64>> caps;          // [round][ncap*4]
  std::vector<u64> final_poly;                 // [len][2]
  u64 pow_witness = 0, pow_response = 0;
  std::vector<u64> indices;                    // [Q]
  std::vector<u64> leaf_len, depth;            // per oracle
  std::vector<std::vector<u64>> init_rows, init_sibs;
  std::vector<u64> step_depth;
  std::vector<std::vector<u64>> step_evals, step_sibs;
  u64 alpha[2] = {0, 0};
  std::vector<u64> betas;                      // [round][2]
  // the polynomial that enters FRI stays on the device ([2][n] columns); copied out only when a test asks for it
  p2b_ctx* ctx = nullptr;
  u64* d_final_in = nullptr;
  u64 final_in_len = 0;
};

extern "C" void p2b_fri_proof_destroy(p2b_fri_proof* p);

// coset LDE of an extension polynomial: coeffs [2][2^k] -> rows [2^(k+rate_bits)][2] (bit-reversed order), on shift * H
static int fri_ext_lde(p2b_ctx* c, const u64* coeffs, u32 k, u32 rate_bits, u64 shift, u64* rows) {
  const u32 log_N = k + rate_bits;
  P2B_TRY(ensure_twiddles(c, log_N > 0 ? log_N - 1 : 0));
  P2B_TRY(ensure_scratch(c, (u64)2 << k));
  ntt::LevelScale sc = lde_scale(k, shift);
  for (u64 b = 0; b < ((u64)1 << rate_bits); b++)
    P2B_TRY(run_lde_block(c, c->stream, coeffs, (u64)1 << k, c->scratch, k, 2, b, sc, rows, 2, 0, b << k));
  return P2B_OK;
}

static int fri_challenger_step(p2b_ctx* c, fri::Challenger* d_ch, const u64* obs, u32 n_obs, u64 obs_cs, u64* out, u32 n_out,
                               u64 modulus) {
  fri::challenger_step_kernel<<<1, 32, 0, c->stream>>>(d_ch, obs, n_obs, obs_cs, out, n_out, modulus);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

// Polynomials [p0, p0 + cnt) of the batch at `point`: enqueues the two kernels on the context's stream and hands back the
// device buffer that will hold the cnt x 2 results (eval_openings_collect copies them out and releases it).  Split in two so
// that a multi-device caller can start every device before it waits for the first (mgpu.cuh).
static int eval_openings_enqueue(p2b_ctx* c, const p2b_batch* b, const uint64_t point[2], u64 p0, u64 cnt, u64** d_out_ret) {
  *d_out_ret = nullptr;
  if (cnt == 0) return P2B_OK;
  Stage stage(c, c->stream, "construct the opening set");   // plonk/prover.rs:208-222
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const u64 n = (u64)1 << b->info.degree_log;
  const u64 seg = 256 * 64;
  const u32 nblk = (u32)((n + seg - 1) / seg);
  hostf::E2h z{point[0] % gl::P, point[1] % gl::P};
  hostf::E2h zbd = hostf::epow(z, 256);
  fri::E2* partial = nullptr;
  u64* d_out = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&partial, cnt * nblk * sizeof(fri::E2), st));
    CUDA_TRY(pool_alloc(&d_out, cnt * 2 * sizeof(u64), st));
    fri::eval_partial_kernel<<<dim3(nblk, (unsigned)cnt), 256, 0, st>>>(b->coeffs + p0 * n, n, seg, fri::E2{z.a, z.b}, fri::E2{zbd.a, zbd.b}, partial);
    fri::eval_finish_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(partial, nblk, cnt, d_out);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return P2B_OK;
  };
  int rc = body();
  if (partial) cudaFreeAsync(partial, st);
  if (rc != P2B_OK) {
    if (d_out) cudaFreeAsync(d_out, st);
    return rc;
  }
  *d_out_ret = d_out;
  return P2B_OK;
}
static int eval_openings_collect(p2b_ctx* c, u64* d_out, u64 cnt, uint64_t* out) {
  if (!d_out) return P2B_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaError_t e = cudaMemcpyAsync(out, d_out, cnt * 2 * sizeof(u64), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFreeAsync(d_out, c->stream);
  CUDA_TRY(e);
  return P2B_OK;
}
extern "C" int p2b_eval_openings(p2b_ctx* c, const p2b_batch* b, const uint64_t point[2], uint64_t* out) {
  if (!c || !b || !point || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  if (!b->coeffs) return fail(P2B_ERR_INVALID, "batch holds no coefficients");
  u64* d_out = nullptr;
  P2B_TRY(eval_openings_enqueue(c, b, point, 0, b->info.num_polys, &d_out));
  return eval_openings_collect(c, d_out, b->info.num_polys, out);
}

// `opener` (multi-device provers, mgpu.cuh): when set, the oracles may be shards -- only their coefficient copies are read here
// -- and the FriInitialTreeProof rows + Merkle paths of oracle o come from opener(o, indices, Q, rows_out, siblings_out), which
// asks the devices that own the leaves.
typedef std::function<int(u32, const u64*, u64, u64*, u64*)> fri_opener;
static int fri_prove_impl(p2b_ctx* c, const p2b_batch* const* oracles, uint32_t num_oracles, const p2b_fri_batch_info* batches,
                          uint32_t num_batches, p2b_challenger* challenger, const p2b_fri_params* params, const fri_opener* opener,
                          p2b_fri_proof** out) {
  if (!c || !oracles || !batches || !challenger || !params || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (num_oracles == 0 || num_batches == 0) return fail(P2B_ERR_INVALID, "no oracles / no opening batches");
  Stage stage(c, c->stream, "compute opening proofs");   // plonk/prover.rs:224-236
  if (challenger->input_len >= 8 || challenger->output_len > 8) return fail(P2B_ERR_INVALID, "challenger buffers out of range");
  if (params->num_reductions && !params->reduction_arity_bits) return fail(P2B_ERR_INVALID, "NULL reduction_arity_bits");
  const u32 k = params->degree_bits, rate_bits = params->rate_bits, cap_height = params->cap_height;
  if (k + rate_bits > 32) return fail(P2B_ERR_INVALID, "degree_bits + rate_bits exceeds the field's two-adicity 32");
  const u64 n = (u64)1 << k, N = n << rate_bits;
  for (u32 i = 0; i < num_oracles; i++) {
    const p2b_batch* o = oracles[i];
    if (!o || !o->coeffs) return fail(P2B_ERR_INVALID, "oracle %u is NULL or holds no coefficients", i);
    if (o->ctx != c) return fail(P2B_ERR_INVALID, "oracle %u belongs to another context", i);
    if (opener) continue;
    if (o->info.degree_log != k || o->info.rate_bits != rate_bits)
      return fail(P2B_ERR_INVALID, "oracle %u: degree_log %u / rate_bits %u do not match the FRI parameters (%u / %u)", i,
                  o->info.degree_log, o->info.rate_bits, k, rate_bits);
    if (o->local_leaves != o->info.num_leaves)
      return fail(P2B_ERR_UNSUPPORTED, "oracle %u is a shard: use p2b_mgpu_fri_prove_openings", i);
  }
  u64 total_arity_bits = 0;
  for (u32 r = 0; r < params->num_reductions; r++) {
    u32 ab = params->reduction_arity_bits[r];
    if (ab == 0 || ab > 8) return fail(P2B_ERR_INVALID, "reduction arity bits %u out of range [1, 8]", ab);
    total_arity_bits += ab;
    // MerkleTree::new's assertion (merkle_tree.rs:285-290) on the tree of this reduction
    u64 leaves_log = k + rate_bits - total_arity_bits;
    if (total_arity_bits > k) return fail(P2B_ERR_INVALID, "reductions fold below degree 1");
    if (cap_height > leaves_log)
      return fail(P2B_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%llu", cap_height, (unsigned long long)leaves_log);
  }
  u64 max_polys = 0;
  for (u32 bi = 0; bi < num_batches; bi++) {
    if (batches[bi].num_polynomials && !batches[bi].polynomials) return fail(P2B_ERR_INVALID, "NULL polynomial list");
    for (u32 j = 0; j < batches[bi].num_polynomials; j++) {
      const p2b_fri_poly_info& pi = batches[bi].polynomials[j];
      if (pi.oracle_index >= num_oracles || pi.polynomial_index >= oracles[pi.oracle_index]->info.num_polys)
        return fail(P2B_ERR_INVALID, "batch %u polynomial %u: (%u, %u) out of range", bi, j, pi.oracle_index, pi.polynomial_index);
    }
    max_polys = std::max<u64>(max_polys, batches[bi].num_polynomials);
  }
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const u32 Q = params->num_query_rounds, R = params->num_reductions;
  const u64 ncap = (u64)1 << cap_height;

  p2b_fri_proof* pr = new (std::nothrow) p2b_fri_proof();
  if (!pr) return fail(P2B_ERR_OOM, "host allocation failed");
  pr->info = p2b_fri_proof_info{R, Q, num_oracles, cap_height, n >> total_arity_bits, N};
  pr->arity_bits.assign(params->reduction_arity_bits, params->reduction_arity_bits + R);

  // device state
  std::vector<void*> to_free;
  auto dalloc = [&](void** p, size_t bytes) -> int {
    CUDA_TRY(pool_alloc(p, bytes ? bytes : 8, st));
    to_free.push_back(*p);
    return P2B_OK;
  };
  std::vector<p2b_batch*> trees;  // commit-phase trees as batches without coefficients (reuses the row/path gather)
  auto body = [&]() -> int {
    fri::Challenger* d_ch = nullptr;
    u64 *d_small = nullptr, *comp = nullptr, *fin = nullptr;
    P2B_TRY(dalloc((void**)&d_ch, sizeof(fri::Challenger)));
    CUDA_TRY(cudaMemcpyAsync(d_ch, challenger, sizeof(fri::Challenger), cudaMemcpyHostToDevice, st));
    // d_small: alpha [2] | betas [R][2] | pow found [1] | pow response [1] | query indices [Q]
    const u64 off_beta = 2, off_found = off_beta + 2 * (u64)R, off_resp = off_found + 1, off_idx = off_resp + 1;
    P2B_TRY(dalloc((void**)&d_small, (off_idx + Q) * sizeof(u64)));
    P2B_TRY(fri_challenger_step(c, d_ch, nullptr, 0, 0, d_small, 2, 0));  // alpha (oracle.rs:1055)

    // ---- final polynomial (oracle.rs:1060-1084) ----
    P2B_TRY(dalloc((void**)&comp, 2 * n * sizeof(u64)));
    CUDA_TRY(pool_alloc(&fin, 2 * n * sizeof(u64), st));
    pr->ctx = c;
    pr->d_final_in = fin;
    pr->final_in_len = n;
    // exponent list: per batch the alpha-power table 0..count-1, then one weight per batch
    std::vector<u64> exps;
    std::vector<u64> table_off(num_batches), weight_exp(num_batches, 0);
    for (u32 bi = 0; bi < num_batches; bi++) {
      table_off[bi] = exps.size();
      for (u32 j = 0; j < batches[bi].num_polynomials; j++) exps.push_back(j);
    }
    // final = sum_b q_b * alpha^(sum of the polynomial counts of the LATER batches): shift_poly multiplies what has
    // been accumulated so far by alpha^count of the batch being added (reducing.rs:108-111, oracle.rs:1080-1081)
    for (u32 bi = 0; bi < num_batches; bi++)
      for (u32 bj = bi + 1; bj < num_batches; bj++) weight_exp[bi] += batches[bj].num_polynomials;
    const u64 weights_off = exps.size();
    for (u32 bi = 0; bi < num_batches; bi++) exps.push_back(weight_exp[bi]);
    u64* d_exps = nullptr;
    fri::E2* d_pows = nullptr;
    P2B_TRY(dalloc((void**)&d_exps, exps.size() * sizeof(u64)));
    P2B_TRY(dalloc((void**)&d_pows, exps.size() * sizeof(fri::E2)));
    CUDA_TRY(cudaMemcpyAsync(d_exps, exps.data(), exps.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    fri::ext_pows_kernel<<<(unsigned)((exps.size() + 127) / 128), 128, 0, st>>>(d_small, d_exps, d_pows, exps.size());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<const u64*> cols;
    std::vector<u64> cols_off(num_batches);
    for (u32 bi = 0; bi < num_batches; bi++) {
      cols_off[bi] = cols.size();
      for (u32 j = 0; j < batches[bi].num_polynomials; j++) {
        const p2b_fri_poly_info& pi = batches[bi].polynomials[j];
        cols.push_back(oracles[pi.oracle_index]->coeffs + (u64)pi.polynomial_index * n);
      }
    }
    const u64** d_cols = nullptr;
    P2B_TRY(dalloc((void**)&d_cols, cols.size() * sizeof(u64*)));
    CUDA_TRY(cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(u64*), cudaMemcpyHostToDevice, st));
    const u32 ch = 64;
    const u64 T = (n + ch - 1) / ch;
    fri::E2* d_tot = nullptr;
    P2B_TRY(dalloc((void**)&d_tot, T * sizeof(fri::E2)));
    for (u32 bi = 0; bi < num_batches; bi++) {
      fri::reduce_polys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cols + cols_off[bi], d_pows + table_off[bi],
                                                                            batches[bi].num_polynomials, n, comp);
      hostf::E2h z{batches[bi].point[0] % gl::P, batches[bi].point[1] % gl::P};
      hostf::E2h zc = hostf::epow(z, ch);
      fri::E2 zd{z.a, z.b};
      fri::scan_chunk_totals_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(comp, n, ch, zd, d_tot, T);
      fri::scan_carries_kernel<<<1, 1024, 0, st>>>(d_tot, T, fri::E2{zc.a, zc.b});
      fri::scan_apply_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(comp, n, ch, zd, d_tot, T, d_pows + weights_off + bi,
                                                                         bi != 0, fin);
      c->launches += 4;
      CUDA_TRY(cudaGetLastError());
    }

    // ---- commit phase (oracle.rs:1086-1092, prover.rs:76-120) ----
    // coefficients are kept as their non-zero prefix [2][n_i]; the LDE zero-pads by 2^rate_bits like lde() + coset_fft
    u64* coeffs = fin;
    u32 ki = k;
    u64 shift = hostf::COSET_SHIFT;
    for (u32 r = 0; r < R; r++) {
      const u32 ab = pr->arity_bits[r];
      const u64 Ni = (u64)1 << (ki + rate_bits), nleaves = Ni >> ab, ll = (u64)2 << ab;
      const u32 lg_leaves = ki + rate_bits - ab;
      p2b_batch* t = new (std::nothrow) p2b_batch();
      if (!t) return fail(P2B_ERR_OOM, "host allocation failed");
      trees.push_back(t);
      t->ctx = c;
      const u64 ndig = 2 * (nleaves - ncap);
      t->info = p2b_batch_info{lg_leaves, 0, cap_height, 0, ll, nleaves, ll, ndig};
      t->shape = merkle::make_shape(lg_leaves, cap_height);
      t->first_leaf = 0;
      t->local_leaves = nleaves;
      CUDA_TRY(pool_alloc(&t->leaves, Ni * 2 * sizeof(u64), st));
      CUDA_TRY(pool_alloc(&t->digests, (ndig ? ndig : 1) * 4 * sizeof(u64), st));
      CUDA_TRY(pool_alloc(&t->cap, ncap * 4 * sizeof(u64), st));
      P2B_TRY(fri_ext_lde(c, coeffs, ki, rate_bits, shift, t->leaves));
      P2B_TRY(launch_hash_leaves(c, st, t->leaves, ll, 1, (u32)ll, 0, nleaves, t->shape, t->digests, t->cap));
      P2B_TRY(launch_layers(c, st, t->shape, t->digests, t->cap, 0, nleaves, 0, &t->top_layer));
      // observe_cap, beta (prover.rs:98-101)
      P2B_TRY(fri_challenger_step(c, d_ch, t->cap, (u32)(ncap * 4), 0, d_small + off_beta + 2 * r, 2, 0));
      u64* folded = nullptr;
      const u64 n_next = ((u64)1 << ki) >> ab;
      P2B_TRY(dalloc((void**)&folded, 2 * n_next * sizeof(u64)));
      fri::fold_kernel<<<(unsigned)((n_next + 255) / 256), 256, 0, st>>>(coeffs, (u64)1 << ki, ab, d_small + off_beta + 2 * r, folded);
      c->launches++;
      CUDA_TRY(cudaGetLastError());
      coeffs = folded;
      ki -= ab;
      shift = hostf::pow(shift, (u64)1 << ab);
    }
    // final polynomial: observe (prover.rs:117-118) and keep
    const u64 flen = (u64)1 << ki;
    P2B_TRY(fri_challenger_step(c, d_ch, coeffs, (u32)(2 * flen), flen, nullptr, 0, 0));
    std::vector<u64> fcols(2 * flen);
    CUDA_TRY(cudaMemcpyAsync(fcols.data(), coeffs, 2 * flen * sizeof(u64), cudaMemcpyDeviceToHost, st));

    // ---- proof of work (prover.rs:123-171) ----
    const u32 min_lz = params->proof_of_work_bits;  // + (64 - F::order().bits()) = + 0
    if (min_lz > 64) return fail(P2B_ERR_INVALID, "proof_of_work_bits > 64");
    unsigned long long* d_found = (unsigned long long*)(d_small + off_found);
    CUDA_TRY(cudaMemsetAsync(d_found, 0xff, sizeof(u64), st));
    u64 found = ~(u64)0;
    // candidates are searched in increasing order, one launch per contiguous range sized ~4x the expected number of
    // tries (2^proof_of_work_bits), so the first launch succeeds with probability ~98% and little work is wasted
    const u64 batch = (u64)1 << std::min<u32>(24, std::max<u32>(14, min_lz + 2));
    for (u64 base = 0; found == ~(u64)0; base += batch) {
      if (base >= gl::P - batch) return fail(P2B_ERR_INVALID, "Proof of work failed. This is highly unlikely!");
      fri::pow_search_kernel<<<(unsigned)(batch / 128), 128, 0, st>>>(d_ch, base, batch, min_lz, d_found);
      c->launches++;
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(&found, d_found, sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
    }
    pr->pow_witness = found;
    // observe the witness, squeeze the response (prover.rs:164-168), then the query challenges (:186)
    P2B_TRY(fri_challenger_step(c, d_ch, d_small + off_found, 1, 0, d_small + off_resp, 1, 0));
    P2B_TRY(fri_challenger_step(c, d_ch, nullptr, 0, 0, d_small + off_idx, Q, N));
    std::vector<u64> small(off_idx + Q);
    CUDA_TRY(cudaMemcpyAsync(small.data(), d_small, small.size() * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(challenger, d_ch, sizeof(fri::Challenger), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    pr->alpha[0] = small[0];
    pr->alpha[1] = small[1];
    pr->betas.assign(small.begin() + off_beta, small.begin() + off_beta + 2 * R);
    pr->pow_response = small[off_resp];
    pr->indices.assign(small.begin() + off_idx, small.end());
    u32 resp_lz = pr->pow_response ? (u32)__builtin_clzll(pr->pow_response) : 64u;
    if (resp_lz < min_lz) return fail(P2B_ERR_INVALID, "internal: proof-of-work response does not verify");
    pr->final_poly.resize(2 * flen);
    for (u64 i = 0; i < flen; i++) {
      pr->final_poly[2 * i] = fcols[i];
      pr->final_poly[2 * i + 1] = fcols[flen + i];
    }
    // ---- caps + query rounds (prover.rs:173-260) ----
    pr->caps.resize(R);
    for (u32 r = 0; r < R; r++) {
      pr->caps[r].resize(ncap * 4);
      CUDA_TRY(cudaMemcpyAsync(pr->caps[r].data(), trees[r]->cap, ncap * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    }
    pr->leaf_len.resize(num_oracles);
    pr->depth.resize(num_oracles);
    pr->init_rows.resize(num_oracles);
    pr->init_sibs.resize(num_oracles);
    for (u32 o = 0; o < num_oracles; o++) {
      pr->leaf_len[o] = oracles[o]->info.leaf_len;
      pr->depth[o] = oracles[o]->shape.sub_log;
      pr->init_rows[o].resize((size_t)Q * pr->leaf_len[o]);
      pr->init_sibs[o].resize((size_t)Q * pr->depth[o] * 4);
      if (Q && opener) P2B_TRY((*opener)(o, pr->indices.data(), Q, pr->init_rows[o].data(), pr->init_sibs[o].data()));
      else if (Q) P2B_TRY(open_impl(oracles[o], pr->indices.data(), Q, pr->init_rows[o].data(), pr->init_sibs[o].data()));
    }
    pr->step_depth.resize(R);
    pr->step_evals.resize(R);
    pr->step_sibs.resize(R);
    std::vector<u64> idx(pr->indices);
    for (u32 r = 0; r < R; r++) {
      const u32 ab = pr->arity_bits[r];
      for (auto& x : idx) x >>= ab;  // prover.rs:243-252
      pr->step_depth[r] = trees[r]->shape.sub_log;
      pr->step_evals[r].resize((size_t)Q * ((u64)2 << ab));
      pr->step_sibs[r].resize((size_t)Q * pr->step_depth[r] * 4);
      if (Q) P2B_TRY(open_impl(trees[r], idx.data(), Q, pr->step_evals[r].data(), pr->step_sibs[r].data()));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return P2B_OK;
  };
  int rc = body();
  for (void* p : to_free) cudaFreeAsync(p, st);
  for (p2b_batch* t : trees) batch_free(t);
  if (rc != P2B_OK) {
    cudaStreamSynchronize(st);
    p2b_fri_proof_destroy(pr);
    return rc;
  }
  *out = pr;
  return P2B_OK;
}

extern "C" int p2b_fri_prove_openings(p2b_ctx* c, const p2b_batch* const* oracles, uint32_t num_oracles,
                                      const p2b_fri_batch_info* batches, uint32_t num_batches, p2b_challenger* challenger,
                                      const p2b_fri_params* params, p2b_fri_proof** out) {
  return fri_prove_impl(c, oracles, num_oracles, batches, num_batches, challenger, params, nullptr, out);
}

extern "C" void p2b_fri_proof_destroy(p2b_fri_proof* p) {
  if (!p) return;
  if (p->d_final_in && ctx_alive(p->ctx)) cudaFreeAsync(p->d_final_in, p->ctx->stream);
  delete p;
}
extern "C" int p2b_fri_proof_get_info(const p2b_fri_proof* p, p2b_fri_proof_info* out) {
  if (!p || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = p->info;
  return P2B_OK;
}
static int copy_out(const std::vector<u64>& v, uint64_t* out) {
  if (!out) return fail(P2B_ERR_INVALID, "NULL output");
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(u64));
  return P2B_OK;
}
extern "C" int p2b_fri_proof_get_cap(const p2b_fri_proof* p, uint32_t round, uint64_t* out) {
  if (!p) return fail(P2B_ERR_INVALID, "NULL proof");
  if (round >= p->caps.size()) return fail(P2B_ERR_INVALID, "round %u out of range", round);
  return copy_out(p->caps[round], out);
}
extern "C" int p2b_fri_proof_get_final_poly(const p2b_fri_proof* p, uint64_t* out) {
  if (!p) return fail(P2B_ERR_INVALID, "NULL proof");
  return copy_out(p->final_poly, out);
}
extern "C" int p2b_fri_proof_get_pow_witness(const p2b_fri_proof* p, uint64_t* out) {
  if (!p || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = p->pow_witness;
  return P2B_OK;
}
extern "C" int p2b_fri_proof_get_query_indices(const p2b_fri_proof* p, uint64_t* out) {
  if (!p) return fail(P2B_ERR_INVALID, "NULL proof");
  return copy_out(p->indices, out);
}
extern "C" int p2b_fri_proof_get_initial(const p2b_fri_proof* p, uint32_t oracle, uint64_t* rows_out, uint64_t* siblings_out) {
  if (!p) return fail(P2B_ERR_INVALID, "NULL proof");
  if (oracle >= p->init_rows.size()) return fail(P2B_ERR_INVALID, "oracle %u out of range", oracle);
  if (rows_out) P2B_TRY(copy_out(p->init_rows[oracle], rows_out));
  if (siblings_out) P2B_TRY(copy_out(p->init_sibs[oracle], siblings_out));
  return P2B_OK;
}
extern "C" int p2b_fri_proof_get_step(const p2b_fri_proof* p, uint32_t round, uint64_t* evals_out, uint64_t* siblings_out,
                                      uint32_t* depth_out) {
  if (!p) return fail(P2B_ERR_INVALID, "NULL proof");
  if (round >= p->step_evals.size()) return fail(P2B_ERR_INVALID, "round %u out of range", round);
  if (evals_out) P2B_TRY(copy_out(p->step_evals[round], evals_out));
  if (siblings_out) P2B_TRY(copy_out(p->step_sibs[round], siblings_out));
  if (depth_out) *depth_out = (u32)p->step_depth[round];
  return P2B_OK;
}
extern "C" int p2b_fri_proof_get_debug(const p2b_fri_proof* p, uint32_t what, uint64_t* out) {
  if (!p || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  switch (what) {
    case 0: out[0] = p->alpha[0]; out[1] = p->alpha[1]; return P2B_OK;
    case 1: return p->betas.empty() ? P2B_OK : copy_out(p->betas, out);
    case 2: {
      const u64 n = p->final_in_len;
      std::vector<u64> cols(2 * n);
      CUDA_TRY(cudaMemcpyAsync(cols.data(), p->d_final_in, 2 * n * sizeof(u64), cudaMemcpyDeviceToHost, p->ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
      for (u64 i = 0; i < n; i++) {
        out[2 * i] = cols[i];
        out[2 * i + 1] = cols[n + i];
      }
      return P2B_OK;
    }
    case 3: out[0] = p->pow_response; return P2B_OK;
    default: return fail(P2B_ERR_INVALID, "unknown debug selector %u", what);
  }
}
