// int_peak.cu -- micro-benchmark of the integer issue rates that bound Poseidon/NTT on B200 (sm_100a).
// Measures thread-instructions per clock per SM for IMAD.WIDE.U32, IMAD (lo), IADD3, LOP3, SHF and mixes, by
// timing long unrolled chains of independent register-only instructions with clock64() on every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak tools/int_peak.cu && ./int_peak
// The SASS of each loop body is checked with cuobjdump (see profiles/int_peak_r01.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint32_t u32;
typedef uint64_t u64;

#define ITERS 16384
#define NACC 8

template <int KIND>
__global__ void __launch_bounds__(256) k(u64* out, u32 a0, u32 b0, long long* cycles) {
  u32 a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
  u64 acc[NACC];
  u32 r[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i] = i * 0x9E3779B97F4A7C15ull + threadIdx.x; r[i] = (u32)acc[i]; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i4 = 0; i4 < NACC * 4; i4++) {
      const int i = i4 % NACC;
      if (KIND == 0) {  // IMAD.WIDE.U32 acc += a*b
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
      } else if (KIND == 1) {  // IMAD lo
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
      } else if (KIND == 2) {  // IADD3
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 3) {  // LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 4) {  // IMAD.HI
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
      } else if (KIND == 5) {  // mix 1:1 IMAD.WIDE + IADD3
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 6) {  // mix 1:2 IMAD.WIDE + 2 IADD3
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 7) {  // 64-bit add with carry chain: IADD3 + IADD3.X
        asm volatile("add.cc.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"(acc[(i + 1) % NACC]));
      } else if (KIND == 8) {  // mix 1:1 IMAD lo + IADD3
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 9) {  // IMAD.WIDE with immediate small constant
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, 41, %0; }" : "+l"(acc[i]));
      } else if (KIND == 10) { // SHF (funnel shift)
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));
      } else if (KIND == 11) { // mix 1:1 IMAD.WIDE + IMAD lo
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
      } else if (KIND == 12) { // FFMA reference
        float f = __uint_as_float(r[i]);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
        r[i] = __float_as_uint(f);
      } else if (KIND == 14) { // IMAD.WIDE.U32 without addend (pure 32x32->64)
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
      } else if (KIND == 15) { // pure mul.wide + 1 IADD3
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 16) { // IMAD.WIDE with addend from a different register pair (acc[i] = lo(acc[i]) * b + acc[(i+1)%N])
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %2; }" : "+l"(acc[i]) : "r"(b), "l"(acc[(i + 1) % NACC]));
      } else if (KIND == 17) { // pure mul.wide + 2 ALU
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 18) { // IMAD.WIDE multiplying by an immediate with a 64-bit accumulate from another pair
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %1; mad.wide.u32 %0, lo, 41, %0; }" : "+l"(acc[i]) : "l"(acc[(i + 1) % NACC]));
      } else if (KIND == 13) { // mix: 1 IMAD.WIDE + 3 ALU
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));
      }
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i] + r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) {
    u32 smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cycles[3 * blockIdx.x] = t0;
    cycles[3 * blockIdx.x + 1] = t1;
    cycles[3 * blockIdx.x + 2] = smid;
  }
}

template <int KIND>
void run(const char* name, int instr_per_slot, int blocks_per_sm) {
  int dev; cudaGetDevice(&dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int nsm = p.multiProcessorCount;
  int nb = nsm * blocks_per_sm;
  u64* out; long long* cyc;
  cudaMalloc(&out, (size_t)nb * 256 * 8); cudaMalloc(&cyc, 3 * nb * sizeof(long long));
  k<KIND><<<nb, 256>>>(out, 12345, 6789, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<KIND><<<nb, 256>>>(out, 12345, 6789, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(3 * nb * sizeof(long long));
  cudaMemcpy(h, cyc, 3 * nb * sizeof(long long), cudaMemcpyDeviceToHost);
  // per SM: instructions executed by the blocks that ran there / (last end - first start)
  double sum_rate = 0; int used = 0; double max_cycles = 0;
  for (int sm = 0; sm < 256; sm++) {
    long long lo = 0, hi = 0; int cnt = 0;
    for (int i = 0; i < nb; i++) if (h[3 * i + 2] == sm) {
      if (!cnt || h[3 * i] < lo) lo = h[3 * i];
      if (!cnt || h[3 * i + 1] > hi) hi = h[3 * i + 1];
      cnt++;
    }
    if (!cnt) continue;
    double instr = (double)cnt * 256 * ITERS * NACC * instr_per_slot;
    sum_rate += instr / (double)(hi - lo); used++;
    if ((double)(hi - lo) > max_cycles) max_cycles = (double)(hi - lo);
  }
  printf("%-34s blocks/SM=%d  %7.1f thread-instr/clk/SM  (%.3f ms, SMs used %d, %.2f GHz from slowest SM)\n", name, blocks_per_sm,
         sum_rate / used, ms, used, max_cycles / (ms * 1e6));
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int bps = 8; bps <= 8; bps *= 2) {
    run<0>("IMAD.WIDE.U32", 1, bps);
    run<9>("IMAD.WIDE.U32 imm", 1, bps);
    run<1>("IMAD (lo)", 1, bps);
    run<4>("IMAD.HI.U32", 1, bps);
    run<2>("IADD3", 1, bps);
    run<3>("LOP3", 1, bps);
    run<10>("SHF", 1, bps);
    run<7>("add.cc.u64 (IADD3+IADD3.X)", 2, bps);
    run<12>("FFMA", 1, bps);
    run<5>("1 IMAD.WIDE + 1 IADD3", 2, bps);
    run<6>("1 IMAD.WIDE + 2 ALU", 3, bps);
    run<13>("1 IMAD.WIDE + 3 ALU", 4, bps);
    run<8>("1 IMAD lo + 1 LOP3", 2, bps);
    run<11>("1 IMAD.WIDE + 1 IMAD lo", 2, bps);
    run<14>("mul.wide (no addend)", 1, bps);
    run<16>("IMAD.WIDE addend other pair", 1, bps);
    run<18>("IMAD.WIDE imm, addend other pair", 1, bps);
    run<15>("1 mul.wide + 1 IADD3", 2, bps);
    run<17>("1 mul.wide + 2 ALU", 3, bps);
    printf("\n");
  }
  return 0;
}
