// pipe_overlap.cu -- do IMAD.WIDE (fmaheavy) and DFMA/DADD (fp64) overlap on sm_100a?  Register-only loops; SASS of the
// loop bodies is checked with cuobjdump before trusting a number (profiles/r02_pipe_model.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint32_t u32;
typedef uint64_t u64;
#define ITERS 4096
#define N 8
template <int W, int D, int A>   // per slot: W wides, D dfma, A extra alu
__global__ void __launch_bounds__(256) k(u64* out, u32 b0, double da, double db, long long* cyc) {
  u32 b = b0 + threadIdx.x;
  u32 x[N], y[N];
  double d[N];
#pragma unroll
  for (int i = 0; i < N; i++) { x[i] = threadIdx.x * 2654435761u + i; y[i] = x[i] ^ 0x5bd1e995u; d[i] = (double)(i + threadIdx.x); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (W) {
        // (lo,hi) = x*b ; x = lo ^ hi ^ y  -> both halves live, one LOP3 per wide
        u32 lo, hi;
        asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[i]), "r"(b));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(lo), "r"(hi), "r"(y[i]));
      }
      if (W == 2) {  // IMAD lo instead of wide (control)
      }
      for (int q = 0; q < D; q++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + q) % N]) : "d"(da), "d"(db));
      for (int q = 0; q < A; q++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 1) % N]), "r"(y[(i + 3) % N]));
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += x[i] + y[i] + (u64)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}
template <int D, int A>
__global__ void __launch_bounds__(256) k32(u64* out, u32 b0, double da, double db, long long* cyc) {  // IMAD lo control
  u32 b = b0 + threadIdx.x;
  u32 x[N], y[N];
  double d[N];
#pragma unroll
  for (int i = 0; i < N; i++) { x[i] = threadIdx.x * 2654435761u + i; y[i] = x[i] ^ 0x5bd1e995u; d[i] = (double)(i + threadIdx.x); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(y[i]));
      for (int q = 0; q < D; q++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + q) % N]) : "d"(da), "d"(db));
      for (int q = 0; q < A; q++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 1) % N]), "r"(y[(i + 3) % N]));
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += x[i] + y[i] + (u64)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}
template <class F>
void run(const char* name, F kern, int per_slot) {
  int nb = 148 * 4;  // 4 blocks x 8 warps = 8 warps per SMSP
  u64* out; long long* cyc;
  cudaMalloc(&out, (size_t)nb * 256 * 8); cudaMalloc(&cyc, 2 * nb * sizeof(long long));
  kern<<<nb, 256>>>(out, 12345, 1.0000001, 0.5, cyc);
  kern<<<nb, 256>>>(out, 12345, 1.0000001, 0.5, cyc);
  cudaDeviceSynchronize();
  long long h[2 * 148 * 4];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nb; i++) avg += (double)(h[2 * i + 1] - h[2 * i]);
  avg /= nb;
  // cycles per slot per warp per SMSP: each SMSP runs 8 warps concurrently (4 blocks x 8 warps / 4 SMSPs)
  double per_slot_cycles = avg / ((double)ITERS * N) / 8.0;
  printf("%-40s %6.2f cycles per slot per warp (%d instr/slot -> %.2f cycles/instr)\n", name, per_slot_cycles, per_slot, per_slot_cycles / per_slot);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run("wide+LOP3", k<1, 0, 0>, 2);
  run("wide+LOP3 + 1 DFMA", k<1, 1, 0>, 3);
  run("wide+LOP3 + 2 DFMA", k<1, 2, 0>, 4);
  run("wide+LOP3 + 3 DFMA", k<1, 3, 0>, 5);
  run("1 DFMA", k<0, 1, 0>, 1);
  run("2 DFMA", k<0, 2, 0>, 2);
  run("2 DFMA + 2 LOP3", k<0, 2, 2>, 4);
  run("wide+LOP3 + 2 LOP3", k<1, 0, 2>, 4);
  run("wide+LOP3 + 2 DFMA + 2 LOP3", k<1, 2, 2>, 6);
  run("IMADlo", k32<0, 0>, 1);
  run("IMADlo + 1 DFMA", k32<1, 0>, 2);
  run("IMADlo + 2 DFMA", k32<2, 0>, 3);
  run("IMADlo + 1 DFMA + 1 LOP3", k32<1, 1>, 3);
  run("IMADlo + 2 DFMA + 2 LOP3", k32<2, 2>, 5);
  return 0;
}
