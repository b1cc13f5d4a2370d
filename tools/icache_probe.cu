// icache_probe.cu -- how does multi-warp issue rate depend on loop-body size on sm_100a?
// Each kernel runs a rolled loop whose body is BODY independent-ish integer instructions (alternating LOP3 / IMAD,
// 8 chains), with 8 warps per SMSP.  Prints warp-instructions per clock per SMSP against body size in KB.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint32_t u32;
typedef uint64_t u64;

template <int BODY>
__global__ void __launch_bounds__(256) k(u32* out, u32 a, u32 b, int iters, long long* cyc, int skew) {
  u32 r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 2654435761u + i;
  // optional skew: warps start the loop at different times so they do not share fetches
  if (skew) {
    int w = threadIdx.x >> 5;
    for (int d = 0; d < w * skew; d++) r[d & 7] = r[d & 7] * a + b;
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < BODY; j++) {
      const int i = j & 7;
      if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
      else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 3) & 7]), "r"(b));
    }
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}

template <int BODY>
void run(int skew) {
  int nsm = 148, bps = 8;  // 8 blocks x 8 warps = 64 warps/SM = 16 per SMSP
  int nb = nsm * bps;
  u32* out; long long* cyc;
  cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&cyc, 2 * nb * sizeof(long long));
  int iters = (1 << 22) / BODY;
  k<BODY><<<nb, 256>>>(out, 3, 5, iters, cyc, skew);
  k<BODY><<<nb, 256>>>(out, 3, 5, iters, cyc, skew);
  cudaDeviceSynchronize();
  long long* h = (long long*)malloc(2 * nb * sizeof(long long));
  cudaMemcpy(h, cyc, 2 * nb * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; i++) avg += (double)(h[2 * i + 1] - h[2 * i]); avg /= nb;
  // all 8 blocks of an SM run concurrently: warp-instr per SMSP = 16 warps * iters * BODY
  double ipc = 16.0 * iters * (BODY + 3) / avg;
  printf("body %6d instr (%5.1f KB) skew %3d : %.3f warp-instr/clk/SMSP\n", BODY, BODY * 16 / 1024.0, skew, ipc);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int skew = 0; skew <= 97; skew += 97) {
    run<64>(skew); run<128>(skew); run<256>(skew); run<384>(skew); run<512>(skew); run<768>(skew); run<1024>(skew);
    run<1536>(skew); run<2048>(skew); run<3072>(skew); run<4096>(skew); run<8192>(skew);
  }
  return 0;
}
