// polynomial_batch.hpp -- C++ host-side mirror of the reference's Rust interface for this path, over the C ABI
// (include/plonky2_b200.h).  Same names, argument meaning and error behaviour as
//   PolynomialBatch::{from_values, from_coeffs, get_lde_values}   plonky2/src/fri/oracle.rs:709-731, 911-977, 1007-1018
//   MerkleTree::{get, prove}, MerkleCap                           plonky2/src/hash/merkle_tree.rs:19-39, 383-440
//   compute_quotient_polys                                        plonky2/src/plonk/prover.rs:790-1034
//   all_wires_permutation_partial_products (+ Z-first ordering)   plonky2/src/plonk/prover.rs:702-786, 112-117
//   OpeningSet::new's eval_commitment                             plonky2/src/plonk/proof.rs:313-319
//   Challenger, FriProof, PolynomialBatch::prove_openings         plonky2/src/iop/challenger.rs:15-150, fri/proof.rs, fri/oracle.rs:1046-1110
// The Rust toolchain is absent in the build image, so this header stands where the reference's host code (Rust) would
// call the FFI; errors the reference raises as panics are thrown as std::runtime_error with the library's message.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/plonky2_b200.h"

namespace plonky2_b200 {

using F = uint64_t;                    // GoldilocksField element (canonical u64)
struct HashOut { F elements[4]; };     // plonky2/src/hash/hash_types.rs

inline void check(int rc) {
  if (rc != P2B_OK) throw std::runtime_error(std::string("plonky2_b200: ") + p2b_last_error());
}

class Context {  // replaces CudaInvContext (fri/oracle.rs:75-109)
 public:
  explicit Context(int device = -1) { check(p2b_ctx_create(device, &ctx_)); }
  ~Context() { p2b_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  p2b_ctx* raw() const { return ctx_; }
  void synchronize() { check(p2b_ctx_synchronize(ctx_)); }

 private:
  p2b_ctx* ctx_ = nullptr;
};

struct MerkleCap { std::vector<HashOut> hashes; size_t height() const { size_t h = 0; while ((size_t(1) << h) < hashes.size()) h++; return h; } };
struct MerkleProof { std::vector<HashOut> siblings; };

class PolynomialBatch;

class MerkleTree {  // device-resident view: leaves, digests, cap of one batch
 public:
  explicit MerkleTree(const PolynomialBatch* b) : b_(b) {}
  MerkleCap cap() const;
  std::vector<F> get(uint64_t leaf_index) const;       // merkle_tree.rs:383-389
  MerkleProof prove(uint64_t leaf_index) const;         // merkle_tree.rs:392-440

 private:
  const PolynomialBatch* b_;
};

class PolynomialBatch {
 public:
  static constexpr size_t SALT_SIZE = P2B_SALT_SIZE;  // fri/oracle.rs:41

  // values / coeffs: one polynomial after another ([P][n]).  blinding: pass the 4 salt columns ([4][n << rate_bits]);
  // the reference draws them with F::rand_vec (oracle.rs:998-1002).
  static PolynomialBatch from_values(Context& ctx, const std::vector<std::vector<F>>& values, size_t rate_bits, bool blinding,
                                     size_t cap_height, const std::vector<F>* salt = nullptr) {
    return make(ctx, values, rate_bits, blinding, cap_height, salt, true);
  }
  static PolynomialBatch from_coeffs(Context& ctx, const std::vector<std::vector<F>>& coeffs, size_t rate_bits, bool blinding,
                                     size_t cap_height, const std::vector<F>* salt = nullptr) {
    return make(ctx, coeffs, rate_bits, blinding, cap_height, salt, false);
  }
  PolynomialBatch(PolynomialBatch&& o) noexcept : b_(o.b_), info_(o.info_), merkle_tree(this) { o.b_ = nullptr; }
  ~PolynomialBatch() { if (b_) p2b_batch_destroy(b_); }

  size_t degree_log() const { return info_.degree_log; }
  size_t rate_bits() const { return info_.rate_bits; }
  bool blinding() const { return info_.salt_size != 0; }
  p2b_batch* raw() const { return b_; }
  const p2b_batch_info& info() const { return info_; }

  // polynomials: Vec<PolynomialCoeffs<F>> (oracle.rs:113)
  std::vector<std::vector<F>> polynomials() const {
    size_t n = size_t(1) << info_.degree_log;
    std::vector<F> flat(info_.num_polys * n);
    check(p2b_batch_get_coeffs(b_, flat.data()));
    std::vector<std::vector<F>> out(info_.num_polys);
    for (size_t c = 0; c < info_.num_polys; c++) out[c].assign(flat.begin() + c * n, flat.begin() + (c + 1) * n);
    return out;
  }
  // get_lde_values(index, step) (oracle.rs:1007-1018): salt stripped
  std::vector<F> get_lde_values(uint64_t index, uint64_t step) const {
    std::vector<F> row(info_.num_polys);
    check(p2b_batch_get_lde_values(b_, index, step, row.data()));
    return row;
  }

  MerkleTree merkle_tree;

 private:
  PolynomialBatch(p2b_batch* b) : b_(b), merkle_tree(this) { check(p2b_batch_get_info(b_, &info_)); }
  static PolynomialBatch make(Context& ctx, const std::vector<std::vector<F>>& polys, size_t rate_bits, bool blinding,
                              size_t cap_height, const std::vector<F>* salt, bool is_values) {
    if (polys.empty()) throw std::runtime_error("plonky2_b200: empty batch (no polynomials)");
    size_t n = polys[0].size();
    if (n == 0 || (n & (n - 1))) throw std::runtime_error("plonky2_b200: polynomial length must be a power of two");
    uint32_t n_log = 0;
    while ((size_t(1) << n_log) < n) n_log++;
    std::vector<F> flat(polys.size() * n);
    for (size_t c = 0; c < polys.size(); c++) {
      if (polys[c].size() != n) throw std::runtime_error("Polynomial degrees inconsistent");  // oracle.rs:991
      std::copy(polys[c].begin(), polys[c].end(), flat.begin() + c * n);
    }
    if (blinding && (!salt || salt->size() != SALT_SIZE * (n << rate_bits)))
      throw std::runtime_error("plonky2_b200: blinding needs 4 salt columns of n << rate_bits elements");
    p2b_batch* b = nullptr;
    auto fn = is_values ? p2b_commit_from_values : p2b_commit_from_coeffs;
    check(fn(ctx.raw(), flat.data(), 1, n_log, polys.size(), (uint32_t)rate_bits, (uint32_t)cap_height,
             blinding ? salt->data() : nullptr, 1, &b));
    return PolynomialBatch(b);
  }
  p2b_batch* b_ = nullptr;
  p2b_batch_info info_{};
};

inline MerkleCap MerkleTree::cap() const {
  MerkleCap c;
  c.hashes.resize(size_t(1) << b_->info().cap_height);
  check(p2b_batch_get_cap(b_->raw(), reinterpret_cast<F*>(c.hashes.data())));
  return c;
}
inline std::vector<F> MerkleTree::get(uint64_t leaf_index) const {
  std::vector<F> row(b_->info().leaf_len);
  check(p2b_batch_get_leaves(b_->raw(), leaf_index, 1, row.data()));
  return row;
}
inline MerkleProof MerkleTree::prove(uint64_t leaf_index) const {
  MerkleProof p;
  p.siblings.resize(b_->info().degree_log + b_->info().rate_bits - b_->info().cap_height);
  check(p2b_batch_prove(b_->raw(), &leaf_index, 1, reinterpret_cast<F*>(p.siblings.data())));
  return p;
}

// compute_quotient_polys (prover.rs:790-1034): returns num_challenges coefficient vectors of n * 2^ceil(log2 qdf) entries.
inline std::vector<std::vector<F>> compute_quotient_polys(Context& ctx, const p2b_circuit& circuit, const PolynomialBatch& wires,
                                                          const PolynomialBatch& zs_partial_products,
                                                          const PolynomialBatch& constants_sigmas, const HashOut& public_inputs_hash,
                                                          const std::vector<F>& betas, const std::vector<F>& gammas,
                                                          const std::vector<F>& alphas) {
  uint32_t qdb = 0;
  while ((1u << qdb) < circuit.quotient_degree_factor) qdb++;
  size_t size = size_t(1) << (circuit.degree_bits + qdb), nc = circuit.num_challenges;
  void* d = nullptr;
  check(p2b_malloc(ctx.raw(), nc * size * sizeof(F), &d));
  int rc = p2b_quotient_polys(ctx.raw(), &circuit, wires.raw(), zs_partial_products.raw(), constants_sigmas.raw(),
                              public_inputs_hash.elements, betas.data(), gammas.data(), alphas.data(), nullptr, static_cast<F*>(d));
  std::vector<F> flat(nc * size);
  if (rc == P2B_OK) rc = p2b_memcpy_d2h(ctx.raw(), flat.data(), d, flat.size() * sizeof(F));
  p2b_free(ctx.raw(), d);
  check(rc);
  std::vector<std::vector<F>> out(nc);
  for (size_t c = 0; c < nc; c++) out[c].assign(flat.begin() + c * size, flat.begin() + (c + 1) * size);
  return out;
}

// all_wires_permutation_partial_products + the Z-first ordering (prover.rs:702-786, :112-117).  wires_values / sigma_values:
// device pointers (column-major values on H); the result is a device matrix [num_challenges * ceil(routed / qdf)][n] that
// PolynomialBatch::from_values takes as a device-resident input.
inline F* partial_products_and_zs(Context& ctx, const F* d_wires_values, const F* d_sigma_values, uint32_t degree_bits,
                                  uint32_t quotient_degree_factor, const std::vector<F>& k_is, const std::vector<F>& betas,
                                  const std::vector<F>& gammas) {
  const uint32_t nr = (uint32_t)k_is.size(), nc = (uint32_t)betas.size();
  const size_t K = (nr + quotient_degree_factor - 1) / quotient_degree_factor;
  void* d = nullptr;
  check(p2b_malloc(ctx.raw(), (nc * K ? nc * K : 1) * (sizeof(F) << degree_bits), &d));
  int rc = p2b_partial_products_and_zs(ctx.raw(), d_wires_values, d_sigma_values, degree_bits, nr, quotient_degree_factor, nc, k_is.data(),
                                       betas.data(), gammas.data(), static_cast<F*>(d));
  if (rc != P2B_OK) p2b_free(ctx.raw(), d);
  check(rc);
  return static_cast<F*>(d);  // release with p2b_free
}

struct Ext { F c[2]; };  // QuadraticExtension<GoldilocksField>: c[0] + c[1] X, X^2 = 7 (goldilocks_extensions.rs:14-28)

// eval_commitment (proof.rs:313-319): every polynomial of the batch at an extension point
inline std::vector<Ext> eval_openings(Context& ctx, const PolynomialBatch& batch, const Ext& point) {
  std::vector<Ext> out(batch.info().num_polys);
  check(p2b_eval_openings(ctx.raw(), batch.raw(), point.c, reinterpret_cast<F*>(out.data())));
  return out;
}

struct Challenger : p2b_challenger {  // iop/challenger.rs:15-21; only the state: permutations run on the device
  Challenger() : p2b_challenger{} {}
};

struct FriQueryStep { std::vector<Ext> evals; MerkleProof merkle_proof; };                                  // fri/proof.rs
struct FriQueryRound { std::vector<std::pair<std::vector<F>, MerkleProof>> initial_trees_proof; std::vector<FriQueryStep> steps; };
struct FriProof {
  std::vector<MerkleCap> commit_phase_merkle_caps;
  std::vector<FriQueryRound> query_round_proofs;
  std::vector<Ext> final_poly;
  F pow_witness = 0;
};
struct FriBatchInfo { Ext point; std::vector<p2b_fri_poly_info> polynomials; };                              // fri/structure.rs:34-38

// PolynomialBatch::prove_openings (fri/oracle.rs:1046-1110); `challenger` is advanced like the reference advances it.
inline FriProof prove_openings(Context& ctx, const std::vector<const PolynomialBatch*>& oracles, const std::vector<FriBatchInfo>& instance,
                               Challenger& challenger, uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                               uint32_t proof_of_work_bits, uint32_t num_query_rounds, const std::vector<uint32_t>& reduction_arity_bits) {
  std::vector<const p2b_batch*> raw;
  for (auto* o : oracles) raw.push_back(o->raw());
  std::vector<p2b_fri_batch_info> bi(instance.size());
  for (size_t i = 0; i < instance.size(); i++)
    bi[i] = p2b_fri_batch_info{{instance[i].point.c[0], instance[i].point.c[1]}, instance[i].polynomials.data(),
                               (uint32_t)instance[i].polynomials.size(), 0};
  p2b_fri_params params{degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, (uint32_t)reduction_arity_bits.size(),
                        reduction_arity_bits.data()};
  p2b_fri_proof* h = nullptr;
  check(p2b_fri_prove_openings(ctx.raw(), raw.data(), (uint32_t)raw.size(), bi.data(), (uint32_t)bi.size(), &challenger, &params, &h));
  FriProof out;
  try {
    p2b_fri_proof_info info;
    check(p2b_fri_proof_get_info(h, &info));
    const size_t Q = info.num_query_rounds, ncap = size_t(1) << info.cap_height;
    out.commit_phase_merkle_caps.resize(info.num_reductions);
    for (uint32_t r = 0; r < info.num_reductions; r++) {
      out.commit_phase_merkle_caps[r].hashes.resize(ncap);
      check(p2b_fri_proof_get_cap(h, r, reinterpret_cast<F*>(out.commit_phase_merkle_caps[r].hashes.data())));
    }
    out.final_poly.resize(info.final_poly_len);
    check(p2b_fri_proof_get_final_poly(h, reinterpret_cast<F*>(out.final_poly.data())));
    check(p2b_fri_proof_get_pow_witness(h, &out.pow_witness));
    out.query_round_proofs.resize(Q);
    for (uint32_t o = 0; o < oracles.size(); o++) {
      const size_t ll = oracles[o]->info().leaf_len, depth = degree_bits + rate_bits - oracles[o]->info().cap_height;
      std::vector<F> rows(Q * ll), sibs(Q * depth * 4);
      check(p2b_fri_proof_get_initial(h, o, rows.data(), sibs.data()));
      for (size_t q = 0; q < Q; q++) {
        MerkleProof mp;
        mp.siblings.resize(depth);
        for (size_t l = 0; l < depth; l++)
          for (int w = 0; w < 4; w++) mp.siblings[l].elements[w] = sibs[(q * depth + l) * 4 + w];
        out.query_round_proofs[q].initial_trees_proof.emplace_back(std::vector<F>(rows.begin() + q * ll, rows.begin() + (q + 1) * ll), mp);
      }
    }
    for (uint32_t r = 0; r < info.num_reductions; r++) {
      uint32_t depth = 0;
      check(p2b_fri_proof_get_step(h, r, nullptr, nullptr, &depth));
      const size_t arity = size_t(1) << reduction_arity_bits[r];
      std::vector<F> ev(Q * arity * 2), sibs(Q * depth * 4);
      check(p2b_fri_proof_get_step(h, r, ev.data(), sibs.data(), nullptr));
      for (size_t q = 0; q < Q; q++) {
        FriQueryStep st;
        st.evals.resize(arity);
        for (size_t a = 0; a < arity; a++) st.evals[a] = Ext{{ev[(q * arity + a) * 2], ev[(q * arity + a) * 2 + 1]}};
        st.merkle_proof.siblings.resize(depth);
        for (size_t l = 0; l < depth; l++)
          for (int w = 0; w < 4; w++) st.merkle_proof.siblings[l].elements[w] = sibs[(q * depth + l) * 4 + w];
        out.query_round_proofs[q].steps.push_back(std::move(st));
      }
    }
  } catch (...) {
    p2b_fri_proof_destroy(h);
    throw;
  }
  p2b_fri_proof_destroy(h);
  return out;
}

}  // namespace plonky2_b200
