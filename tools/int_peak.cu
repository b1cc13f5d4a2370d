// int_peak.cu -- micro-benchmark of the integer issue rates that bound Poseidon/NTT on B200 (sm_100a).
// Measures thread-instructions per clock per SM for IMAD.WIDE.U32, IMAD (lo), IADD3, LOP3, SHF and mixes, by
// timing long unrolled chains of independent register-only instructions with clock64() on every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak tools/int_peak.cu && ./int_peak
// The SASS of each loop body is checked with cuobjdump (see profiles/int_peak_r01.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef uint32_t u32;
typedef uint64_t u64;

#define ITERS 16384
#define NACC 8

template <int KIND>
__global__ void __launch_bounds__(256) k(u64* out, u32 a0, u32 b0, long long* cycles) {
  u32 a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
  u64 acc[NACC];
  u32 r[NACC];
  double d[NACC];
  double da = 1.0 + 1e-9 * a0, db = 1e-3 * b0;
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i] = i * 0x9E3779B97F4A7C15ull + threadIdx.x; r[i] = (u32)acc[i]; d[i] = (double)(i + threadIdx.x); }
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
      } else if (KIND == 20) { // DFMA
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
      } else if (KIND == 21) { // DADD
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(db));
      } else if (KIND == 22) { // DMUL
        asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
      } else if (KIND == 23) { // DFMA + IADD3
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 24) { // DFMA + IMAD lo
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
      } else if (KIND == 25) { // DFMA + mul.wide (no addend)
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
      } else if (KIND == 26) { // DFMA + mul.wide + IADD3
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 27) { // DFMA + IMAD lo + IADD3
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i + 4) % NACC]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 28) { // 2 DFMA + mul.wide + 2 ALU
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + 4) % NACC]) : "d"(db), "d"(da));
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i + 4) % NACC]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 29) { // cvt.rn.f64.u32 (I2F.F64.U32)
        asm volatile("{ .reg .f64 t; cvt.rn.f64.u32 t, %0; mov.b64 {%0, _}, t; }" : "+r"(r[i]));
      } else if (KIND == 30) { // cvt.rzi.u32.f64 (F2I)
        asm volatile("{ .reg .u32 t; cvt.rzi.u32.f64 t, %0; mov.b64 %0, {t, t}; }" : "+d"(d[i]));
      } else if (KIND == 31) { // SEL via setp + selp
        asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
      } else if (KIND == 32) { // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0x3715;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]));
      } else if (KIND == 33) { // 32-bit carry chain add.cc / addc (IADD3 + IADD3.X)
        asm volatile("{ add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3; }" : "+r"(r[i]), "+r"(r[(i + 4) % NACC]) : "r"(a), "r"(b));
      } else if (KIND == 34) { // mul.wide + 3 ALU
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));
      } else if (KIND == 35) { // IMAD.WIDE addend other pair + 2 ALU
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %2; }" : "+l"(acc[i]) : "r"(b), "l"(acc[(i + 1) % NACC]));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 36) { // IMAD.WIDE with 32-bit zero-extended addend in own low word, i.e. acc = lo*b + (u64)r
        asm volatile("{ .reg .u32 lo, hi; .reg .u64 c; mov.b64 {lo,hi}, %0; cvt.u64.u32 c, %2; mad.wide.u32 %0, lo, %1, c; }" : "+l"(acc[i]) : "r"(b), "r"(r[i]));
      } else if (KIND == 37) { // DFMA + IMAD.WIDE (addend) 1:1
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[i]) : "r"(b));
      } else if (KIND == 38) { // 3 DFMA + 1 mul.wide + 2 ALU
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + 3) % NACC]) : "d"(db), "d"(da));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + 5) % NACC]) : "d"(db), "d"(da));
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(acc[i]) : "r"(b));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 3) % NACC]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i + 4) % NACC]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 39) { // FFMA + IMAD lo + IADD3 (is fmalite a third issue port?)
        float f = __uint_as_float(r[(i + 2) % NACC]);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
        r[(i + 2) % NACC] = __float_as_uint(f);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i + 4) % NACC]) : "r"(r[(i + 1) % NACC]), "r"(r[(i + 5) % NACC]));
      } else if (KIND == 40) { // mul.wide by immediate (no addend)
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mul.wide.u32 %0, lo, 41; }" : "+l"(acc[i]));
      } else if (KIND == 41) { // IMAD lo with immediate
        asm volatile("mad.lo.u32 %0, %0, 41, %1;" : "+r"(r[i]) : "r"(b));
      } else if (KIND == 42) { // IMAD.SHL style: mul.lo by power of two + add (might go ALU as LEA)
        asm volatile("{ .reg .u32 t; shl.b32 t, %0, 5; add.u32 %0, t, %1; }" : "+r"(r[i]) : "r"(r[(i + 1) % NACC]));
      } else if (KIND >= 100 && KIND < 200) {
        // clean mixes (SASS-verified): KIND = 100 + 16*W + 4*D + A  with W in {0: none, 1: mul.wide(+LOP3 keeping both halves live), 2: IMAD lo}, D = #DFMA (0..3), A = #extra LOP3 (0..3)
        const int W = (KIND - 100) / 16, D = ((KIND - 100) / 4) % 4, A = (KIND - 100) % 4;
        if (W == 1) {
          u32 lo, hi;
          asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t; }" : "=r"(lo), "=r"(hi) : "r"(r[i]), "r"(b));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r[i]) : "r"(lo), "r"(hi), "r"(a));
        } else if (W == 2) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
        }
        for (int q = 0; q < D; q++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[(i + 3 * q) % NACC]) : "d"(da), "d"(db));
        for (int q = 0; q < A; q++) { u32 lo_, hi_; asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo_), "=r"(hi_) : "l"(acc[(i + q) % NACC])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo_) : "r"(hi_), "r"(b)); asm volatile("mov.b64 %0, {%1,%2};" : "=l"(acc[(i + q) % NACC]) : "r"(lo_), "r"(hi_)); }
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
  for (int i = 0; i < NACC; i++) s += acc[i] + r[i] + (u64)__double_as_longlong(d[i]);
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

int main(int argc, char** argv) {
  for (int bps = (argc > 1 ? atoi(argv[1]) : 8); bps <= (argc > 1 ? atoi(argv[1]) : 8); bps *= 2) {
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

    run<20>("DFMA", 1, bps);
    run<21>("DADD", 1, bps);
    run<22>("DMUL", 1, bps);
    run<23>("1 DFMA + 1 IADD3", 2, bps);
    run<24>("1 DFMA + 1 IMAD lo", 2, bps);
    run<25>("1 DFMA + 1 mul.wide", 2, bps);
    run<37>("1 DFMA + 1 IMAD.WIDE(addend)", 2, bps);
    run<26>("1 DFMA + 1 mul.wide + 1 IADD3", 3, bps);
    run<27>("1 DFMA + 1 IMAD lo + 1 LOP3", 3, bps);
    run<28>("2 DFMA + 1 mul.wide + 2 ALU", 5, bps);
    run<38>("3 DFMA + 1 mul.wide + 2 ALU", 6, bps);
    run<39>("1 FFMA + 1 IMAD lo + 1 LOP3", 3, bps);
    run<29>("I2F.F64.U32", 1, bps);
    run<30>("F2I.U32.F64", 1, bps);
    run<31>("ISETP + SEL", 2, bps);
    run<32>("PRMT", 1, bps);
    run<33>("add.cc.u32 + addc.u32", 2, bps);
    run<34>("1 mul.wide + 3 ALU", 4, bps);
    run<35>("1 IMAD.WIDE(addend other) + 2 ALU", 3, bps);
    run<36>("IMAD.WIDE 32-bit addend", 1, bps);
    run<40>("mul.wide imm", 1, bps);
    run<41>("IMAD lo imm", 1, bps);
    run<42>("SHL + IADD (LEA?)", 1, bps);

    printf("-- clean mixes: W=mul.wide(+1 LOP3), I=IMAD lo, D=DFMA, A=LOP3\n");
    run<100 + 16 * 1>("W(+L)", 2, bps);
    run<100 + 16 * 1 + 4 * 1>("W(+L) + 1D", 3, bps);
    run<100 + 16 * 1 + 4 * 2>("W(+L) + 2D", 4, bps);
    run<100 + 16 * 1 + 4 * 3>("W(+L) + 3D", 5, bps);
    run<100 + 16 * 1 + 2>("W(+L) + 2A", 4, bps);
    run<100 + 16 * 1 + 4 * 2 + 2>("W(+L) + 2D + 2A", 6, bps);
    run<100 + 4 * 1>("1D", 1, bps);
    run<100 + 4 * 2>("2D", 2, bps);
    run<100 + 4 * 2 + 2>("2D + 2A", 4, bps);
    run<100 + 4 * 1 + 1>("1D + 1A", 2, bps);
    run<100 + 2>("2A", 2, bps);
    run<100 + 16 * 2>("I", 1, bps);
    run<100 + 16 * 2 + 4 * 1>("I + 1D", 2, bps);
    run<100 + 16 * 2 + 4 * 2>("I + 2D", 3, bps);
    run<100 + 16 * 2 + 1>("I + 1A", 2, bps);
    run<100 + 16 * 2 + 4 * 1 + 1>("I + 1D + 1A", 3, bps);
    run<100 + 16 * 2 + 4 * 2 + 2>("I + 2D + 2A", 5, bps);
    printf("\n");
  }
  return 0;
}
