// poseidon_lab.cu -- stand-alone A/B harness for the Poseidon / leaf-hash kernels (csrc/poseidon.cuh, merkle.cuh).
// Compiles the product's kernels into a small executable (seconds instead of the minutes of the whole library), checks
// the reference's four permutation known-answer vectors (poseidon_goldilocks.rs:289-310), then times
// merkle::hash_leaves_kernel on 2^20 leaves x 135 and prints a checksum of all digests so variants can be compared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DVARIANT...] -o lab tools/poseidon_lab.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../plonky2-gpu_b200/csrc/merkle.cuh"

typedef uint64_t u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static const u64 NEG1 = 0xFFFFFFFF00000000ull;
static const u64 KAT_IN[4][12] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
    {NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1, NEG1},
    {0x8ccbbbea4fe5d2b7, 0xc2af59ee9ec49970, 0x90f7e1a9e658446a, 0xdcc0630a3ab8b1b8, 0x7ff8256bca20588c, 0x5d99a7ca0c44ecfb,
     0x48452b17a70fbee3, 0xeb09d654690b6c88, 0x4a55d3a39c676a88, 0xc0407a38d2285139, 0xa234bac9356386d1, 0xe1633f2bad98a52f}};
static const u64 KAT_OUT[4][12] = {
    {0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca, 0xd7709673896996dc, 0x46a84e87642f44ed,
     0xd032648251ee0b3c, 0x1c687363b207df62, 0xdf8565563e8045fe, 0x40f5b37ff4254dae, 0xd070f637b431067c, 0x1792b1c4342109d7},
    {0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2, 0x414754bfebd051f0, 0xcb1f8980294a023f,
     0x6eb2a9e4d54a9d0f, 0x1902bc3af467e056, 0xf045d5eafdc6021f, 0xe4150f77caaa3be5, 0xc9bfd01d39b50cce, 0x5c0a27fcb0e1459b},
    {0xbe0085cfc57a8357, 0xd95af71847d05c09, 0xcf55a13d33c1c953, 0x95803a74f4530e82, 0xfcd99eb30a135df1, 0xe095905e913a3029,
     0xde0392461b42919b, 0x7d3260e24e81d031, 0x10d3d0465d9deaa0, 0xa87571083dfc2a47, 0xe18263681e9958f8, 0xe28e96f1ae5e60d3},
    {0xa89280105650c4ec, 0xab542d53860d12ed, 0x5704148e9ccab94f, 0xd3a826d4b62da9f5, 0x8a7a6ca87892574f, 0xc7017e1cad1a674e,
     0x1f06668922318e34, 0xa3b203bc8102676f, 0xfcc781b0ce382bf2, 0x934c69ff3ed14ba5, 0x504688a5996e8f13, 0x401f3f2ed524a2ba}};

__device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void fill_kernel(u64* out, u64 count, u64 seed, int noncanonical_every) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 v = splitmix64(seed ^ splitmix64(i));
  while (v >= gl::P) v = splitmix64(v);
  // sprinkle edge values: 0, p-1, 2^32-1, 2^48 (forces the rare reduction case), non-canonical representatives
  if (noncanonical_every && (i % noncanonical_every) == 0) {
    const u64 edge[6] = {0, gl::P - 1, 0xFFFFFFFFull, 1ull << 48, 0xFFFFFFFFFFFFFFFFull, gl::P};
    v = edge[(i / noncanonical_every) % 6];
  }
  out[i] = v;
}
__global__ void checksum_kernel(const u64* d, u64 count, u64* out) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 v = 0;
  for (u64 k = i; k < count; k += (u64)gridDim.x * blockDim.x) v += d[k] * (2 * k + 1);
  atomicAdd((unsigned long long*)out, (unsigned long long)v);
}

int main(int argc, char** argv) {
  int log_leaves = argc > 1 ? atoi(argv[1]) : 20;
  int P = argc > 2 ? atoi(argv[2]) : 135;
  CK(poseidon::upload_constants());
  // ---- KATs through permute_kernel ----
  u64* d_st;
  CK(cudaMalloc(&d_st, 4 * 12 * 8));
  CK(cudaMemcpy(d_st, KAT_IN, sizeof(KAT_IN), cudaMemcpyHostToDevice));
  merkle::permute_kernel<<<1, P2B_HASH_BLOCK>>>(d_st, 4, 1);
  CK(cudaDeviceSynchronize());
  u64 got[4][12];
  CK(cudaMemcpy(got, d_st, sizeof(got), cudaMemcpyDeviceToHost));
  bool kat = true;
  for (int k = 0; k < 4; k++)
    for (int i = 0; i < 12; i++) kat &= got[k][i] == KAT_OUT[k][i];
  // ---- leaf hashing ----
  u64 N = 1ull << log_leaves;
  u64 *d_leaves, *d_dig, *d_cap, *d_sum;
  CK(cudaMalloc(&d_leaves, N * P * 8));
  CK(cudaMalloc(&d_dig, 2 * N * 32));
  CK(cudaMalloc(&d_cap, 16 * 32));
  CK(cudaMalloc(&d_sum, 8));
  fill_kernel<<<(unsigned)((N * P + 255) / 256), 256>>>(d_leaves, N * P, 7, 1000003);
  merkle::TreeShape shape = merkle::make_shape(log_leaves, 4);
  unsigned grid = (unsigned)((N + P2B_HASH_BLOCK - 1) / P2B_HASH_BLOCK);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    merkle::hash_leaves_kernel<<<grid, P2B_HASH_BLOCK>>>(d_leaves, (u64)P, 1, (unsigned)P, N, 0, shape, d_dig, d_cap);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  CK(cudaMemset(d_sum, 0, 8));
  checksum_kernel<<<1024, 256>>>(d_dig, shape.sub_digests * 16 * 4, d_sum);  // leaf digests live at their layout slots; the rest is whatever cudaMalloc left
  // only leaf-digest slots are deterministic: checksum them explicitly instead
  CK(cudaDeviceSynchronize());
  // deterministic checksum: re-hash into a zeroed buffer
  CK(cudaMemset(d_dig, 0, 2 * N * 32));
  merkle::hash_leaves_kernel<<<grid, P2B_HASH_BLOCK>>>(d_leaves, (u64)P, 1, (unsigned)P, N, 0, shape, d_dig, d_cap);
  CK(cudaMemset(d_sum, 0, 8));
  checksum_kernel<<<1024, 256>>>(d_dig, 2 * N * 4, d_sum);
  u64 sum;
  CK(cudaMemcpy(&sum, d_sum, 8, cudaMemcpyDeviceToHost));
  double perms = (double)N * ((P + 7) / 8);
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, merkle::hash_leaves_kernel));
  printf("%-28s KAT %s  hash 2^%d x %d: %8.3f ms  %7.1f Mperm/s  checksum %016llx  regs %d  local %zu B\n",
#ifdef LAB_NAME
         LAB_NAME,
#else
         "default",
#endif
         kat ? "ok" : "FAIL", log_leaves, P, best, perms / best / 1e3, (unsigned long long)sum, fa.numRegs, fa.localSizeBytes);
  return kat ? 0 : 2;
}
