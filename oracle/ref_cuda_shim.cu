// TEST INFRASTRUCTURE ONLY.  Compiles the reference's own CUDA translation unit UNMODIFIED, from the source tree where
// it lies (the path is passed as -DP2REF_TU=...; nothing is copied into this repository), into
// oracle/_ref/libplonky2_ref_cuda.so.  tests/test_ref_cuda_crosscheck.py runs its `ifft` and `merkle_tree_from_coeffs`
// (cuda/plonky2_gpu.cu:70-86, 435-606) on the GPU box on the same inputs as the oracle and as this library's
// reference-compatible entry points: that pins the oracle's LDE values / digests / cap to outputs of the reference
// itself (its GPU path; the Rust CPU path cannot be built in this image).  Only rate_bits = 3 and n >= 512 are
// usable: init_lde_kernel hard-codes 7 = 2^3 - 1 (plonky2_gpu_impl.cuh:290-294) and ifft_kernel asserts
// perpoly_thcnt < values_num_per_poly with 256 threads per polynomial.
#ifndef P2REF_TU
#error "pass -DP2REF_TU=\"/root/reference/cuda/plonky2_gpu.cu\""
#endif
#include P2REF_TU

// ---- gate-constraint cross-check --------------------------------------------------------------------------------
// The reference's OWN device-side gate evaluators (cuda/*Gate.cuh: eval_unfiltered_base_packed, instantiated the way
// cuda/plonky2_gpu_impl.cuh:633-685 instantiates them) run on caller-supplied rows; one thread per row writes the
// gate's constraint values.  tests/test_ref_cuda_crosscheck.py compares them with oracle/quotient.py's restatement,
// which pins the oracle's gate constraints to the reference's code.  Nothing of the reference is copied here: this file
// only calls what `#include P2REF_TU` brought in.
//   type ids = include/plonky2_b200.h P2B_GATE_*; p0..p2 = the gate's parameters in the same slots.
template <class G>
__device__ void p2ref_run_gate(G g, GoldilocksField* consts, int ncst, GoldilocksField* wires, int nw, const uint64_t* pih,
                               GoldilocksField* out, int nout) {
  PoseidonHasher::HashOut h;
  for (int i = 0; i < 4; i++) h.elements[i] = GoldilocksField::from_canonical_u64(pih[i]);
  EvaluationVarsBasePacked vars{GoldilocksFieldView{consts, ncst}, GoldilocksFieldView{wires, nw}, h, 0};
  g.eval_unfiltered_base_packed(vars, StridedConstraintConsumer{out, out + nout});
}

__global__ void p2ref_gate_kernel(int type, int p0, int p1, int p2, uint64_t* wires, int nw, uint64_t* consts, int ncst,
                                  const uint64_t* pih, uint64_t* out, int nout, int rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  GoldilocksField* w = (GoldilocksField*)(wires + (size_t)r * nw);
  GoldilocksField* k = (GoldilocksField*)(consts + (size_t)r * ncst);
  GoldilocksField* o = (GoldilocksField*)(out + (size_t)r * nout);
  switch (type) {
    case 0: { NoopGate g; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 1: { ConstantGate g{.num_consts = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 2: { PublicInputGate g; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 3: { ArithmeticGate g{.num_ops = p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 4:
      if (p1 == 2) { BaseSumGate<2> g{.num_limbs = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); }
      else { BaseSumGate<4> g{.num_limbs = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); }
      break;
    case 5: { PoseidonGate g; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 6: { RandomAccessGate g{.bits = (usize)p0, .num_copies = (usize)p1, .num_extra_constants = (usize)p2}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 7: { U32ArithmeticGate g{.num_ops = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 8: { U32AddManyGate g{.num_addends = (usize)p0, .num_ops = (usize)p1}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 9: { U32RangeCheckGate g{.num_input_limbs = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 10: { U32SubtractionGate g{.num_ops = (usize)p0}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    case 11: { ComparisonGate g{.num_bits = (usize)p0, .num_chunks = (usize)p1}; p2ref_run_gate(g, k, ncst, w, nw, pih, o, nout); break; }
    default: break;
  }
}

// host entry: device pointers wires [rows][nw], consts [rows][ncst] (selector prefix already removed), pih [4],
// out [rows][nout].  Returns the CUDA error code.
extern "C" int p2ref_eval_gate(int type, int p0, int p1, int p2, uint64_t* d_wires, int nw, uint64_t* d_consts, int ncst,
                               const uint64_t* d_pih, uint64_t* d_out, int nout, int rows) {
  p2ref_gate_kernel<<<(rows + 63) / 64, 64>>>(type, p0, p1, p2, d_wires, nw, d_consts, ncst, d_pih, d_out, nout, rows);
  cudaError_t e = cudaDeviceSynchronize();
  return (int)e;
}

// ---- quotient values: the reference's own kernel on caller-supplied leaves -------------------------------------------
// compute_quotient_values_kernel (cuda/plonky2_gpu_impl.cuh:485-876) is compiled for ONE circuit (25 gate instances in 6
// selector groups, 2 challenges, 80 routed wires, quotient degree factor 8, 231 gate constraints, rate_bits 3).  This
// entry launches it on caller-supplied device data; the test builds the same circuit for the oracle and for
// libplonky2_b200 and compares all three.  All pointers are device pointers except pih (host, 4 words).
extern "C" int p2ref_quotient_values(int degree_log, uint64_t* d_points, uint64_t* d_outs, const uint64_t* pih,
                                     uint64_t* d_cs_leaves, int cs_leaf_len, uint64_t* d_zs_leaves, int zs_leaf_len,
                                     uint64_t* d_wires_leaves, int wires_leaf_len, int num_constants, int num_partial_products,
                                     uint64_t* d_zh, uint64_t* d_zh_inv, uint64_t* d_k_is, uint64_t* d_alphas, uint64_t* d_betas,
                                     uint64_t* d_gammas) {
  PoseidonHasher::HashOut h;
  for (int i = 0; i < 4; i++) h.elements[i] = GoldilocksField{pih[i]};
  const int lde = 1 << (degree_log + 3);
  compute_quotient_values_kernel<<<(lde + 31) / 32, 32>>>(
      degree_log, 3, (GoldilocksField*)d_points, (GoldilocksField*)d_outs, h, (GoldilocksField*)d_cs_leaves, cs_leaf_len,
      (GoldilocksField*)d_zs_leaves, zs_leaf_len, (GoldilocksField*)d_wires_leaves, wires_leaf_len, num_constants, 80, 2, 231, 8,
      num_partial_products, (GoldilocksField*)d_zh, (GoldilocksField*)d_zh_inv, (GoldilocksField*)d_k_is, (GoldilocksField*)d_alphas,
      (GoldilocksField*)d_betas, (GoldilocksField*)d_gammas);
  cudaError_t e = cudaDeviceSynchronize();
  return (int)e;
}
