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
