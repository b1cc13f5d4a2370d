// polynomial_batch.hpp -- C++ host-side mirror of the reference's Rust interface for this path, over the C ABI
// (include/plonky2_b200.h).  Same names, argument meaning and error behaviour as
//   PolynomialBatch::{from_values, from_coeffs, get_lde_values}   plonky2/src/fri/oracle.rs:709-731, 911-977, 1007-1018
//   MerkleTree::{get, prove}, MerkleCap                           plonky2/src/hash/merkle_tree.rs:19-39, 383-440
//   compute_quotient_polys                                        plonky2/src/plonk/prover.rs:790-1034
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

}  // namespace plonky2_b200
