// mds_fft.cuh -- the Poseidon MDS layer (circulant [17,15,41,16,2,28,13,13,39,18,34,20] + diag(8,0,...),
// plonky2/src/hash/poseidon_goldilocks.rs:21-22; semantics of mds_row_shf / mds_layer, poseidon.rs:172-260) as a
// length-12 cyclic convolution over the integers, split by the Chinese remainder theorem
//     Z[t]/(t^12 - 1)  =  Z[t]/(t^3 - 1)  x  Z[t]/(t^3 + 1)  x  Z[i][t]/(t^3 - i)   (+ the conjugate of the last).
// For THIS matrix every transformed coefficient is a signed power of two:
//     K(+1)/4 = [16, 32, 16],   K(-1)/4 = [-1, -8, 2],   K(i)/2 = [2+i, -4-i, 16-i]
// so one layer on a vector of 12 integers is 78 additions / multiply-adds by +-2^k and no general multiplication
// (the direct form is 145 multiply-accumulates).  The arithmetic is exact ring arithmetic: T may be
//   * double   -- the FP64 pipe of sm_100a (64 DFMA lanes/clk/SM, idle in an integer kernel): inputs are the 32-bit
//                 halves of the state words, every intermediate is an integer of magnitude < 2^32 * 264 * 4 < 2^53,
//   * int64_t / uint32_t -- two's-complement integers (wrap-around is harmless: the true result fits).
// Derivation and the numerical check against the direct form: tools/mds_fft_check.cpp (built by tests/test_mds_fft_cpu.py).
#pragma once
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace mdsfft {

// y = x*k + a for a small signed power-of-two k
template <class T>
__host__ __device__ __forceinline__ T mad(T x, int k, T a) { return x * (T)k + a; }
#ifdef __CUDACC__
template <>
__device__ __forceinline__ double mad<double>(double x, int k, double a) { return fma(x, (double)k, a); }
#endif

// y[r] = sum_i circ[i] * x[(i + r) % 12] + 8 * x[0] * [r == 0]
template <class T>
__host__ __device__ __forceinline__ void mds12(const T (&x)[12], T (&y)[12]) {
  T X1[3], Xm[3], R[3], I[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    T e = x[j] + x[j + 6], o = x[j + 3] + x[j + 9];
    R[j] = x[j] - x[j + 6];
    I[j] = x[j + 3] - x[j + 9];
    X1[j] = e + o;
    Xm[j] = e - o;
  }
  // t^3 = 1 component (x 1/16): cyclic convolution with [1, 2, 1]
  T S = (X1[0] + X1[1]) + X1[2];
  T A[3] = {S + X1[2], S + X1[0], S + X1[1]};
  // t^3 = -1 component: negacyclic convolution with [-1, -8, 2]
  T B[3];
  B[0] = mad(Xm[2], 8, mad(Xm[1], -2, -Xm[0]));
  B[1] = mad(Xm[0], -8, mad(Xm[2], -2, -Xm[1]));
  B[2] = mad(Xm[1], -8, mad(Xm[0], 2, -Xm[2]));
  // t^3 = i component: (R + iI) * [2+i, -4-i, 16-i]
  T re[3], im[3];
  re[0] = mad(I[1], -16, mad(I[2], 4, mad(R[0], 2, (R[1] + R[2]) - I[0])));
  im[0] = mad(R[1], 16, mad(R[2], -4, mad(I[0], 2, (I[1] + I[2]) + R[0])));
  re[1] = mad(I[2], -16, mad(R[0], -4, mad(R[1], 2, (I[0] + R[2]) - I[1])));
  im[1] = mad(R[2], 16, mad(I[0], -4, mad(I[1], 2, (R[1] + I[2]) - R[0])));
  re[2] = mad(R[0], 16, mad(R[1], -4, mad(R[2], 2, (I[0] + I[1]) - I[2])));
  im[2] = mad(I[0], 16, mad(I[1], -4, mad(I[2], 2, (R[2] - R[0]) - R[1])));
  // recombine: y[j + 3m] = 16 A[j] + (-1)^m B[j] + Re(i^-m (re + i im))
#pragma unroll
  for (int j = 0; j < 3; j++) {
    T P = mad(A[j], 16, B[j]), Q = mad(A[j], 16, -B[j]);
    y[j] = P + re[j];
    y[j + 6] = P - re[j];
    y[j + 3] = Q + im[j];
    y[j + 9] = Q - im[j];
  }
  y[0] = mad(x[0], 8, y[0]);
}

}  // namespace mdsfft
