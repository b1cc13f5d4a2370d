// FRI opening proof on the device: kernels.
//
// Reference behaviour restated (nothing is copied; the reference runs all of this on the CPU):
//   quadratic extension F[X]/(X^2-7)     field/src/extension/quadratic.rs:172-185, goldilocks_extensions.rs:14-28
//   Challenger (overwrite-mode duplex)   plonky2/src/iop/challenger.rs:15-150
//   OpeningSet evaluation                plonky2/src/plonk/proof.rs:305-334
//   reduce_polys_base / divide_by_linear plonky2/src/util/reducing.rs:87-111, field/src/polynomial/division.rs:73-88
//   commit-phase fold                    plonky2/src/fri/prover.rs:76-120
//   proof-of-work grinding               plonky2/src/fri/prover.rs:123-171
//
// Layouts: an extension polynomial's coefficients are two base-field columns [2][len] (c0 column, c1 column), which is
// what the NTT kernels take as a 2-polynomial batch; its evaluations come out of ntt_final_pass_kernel<MODE_ROWS> as rows
// [len][2] in bit-reversed order, which IS the flattened leaf matrix [len/arity][2*arity] of a commit-phase tree
// (prover.rs:90-96: reverse_index_bits, chunks(arity), flatten).
#pragma once
#include "gl64.cuh"
#include "poseidon.cuh"

namespace fri {
using gl::u64;
typedef uint32_t u32;

struct E2 {
  u64 a, b;  // a + b*X
};
__device__ __forceinline__ E2 eadd(E2 x, E2 y) { return E2{gl::add(x.a, y.a), gl::add(x.b, y.b)}; }
__device__ __forceinline__ E2 esub(E2 x, E2 y) { return E2{gl::sub(x.a, y.a), gl::sub(x.b, y.b)}; }
// (a0 + a1 X)(b0 + b1 X) = a0 b0 + 7 a1 b1 + (a0 b1 + a1 b0) X      quadratic.rs:176-184
__device__ __forceinline__ E2 emul(E2 x, E2 y) {
  u64 t = gl::mul(x.b, y.b);
  return E2{gl::mul_add(x.a, y.a, gl::mul(7, t)), gl::mul_add(x.a, y.b, gl::mul(x.b, y.a))};
}
// acc * z + c with c in the base field (Horner step of eval / divide_by_linear on base coefficients)
__device__ __forceinline__ E2 emul_add_base(E2 acc, E2 z, u64 c) {
  E2 r = emul(acc, z);
  r.a = gl::add(r.a, c);
  return r;
}
__device__ __forceinline__ E2 epow(E2 x, u64 e) {
  E2 r{1, 0};
  while (e) {
    if (e & 1) r = emul(r, x);
    x = emul(x, x);
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ E2 ecanon(E2 x) { return E2{gl::canon(x.a), gl::canon(x.b)}; }

// ---- Challenger ---------------------------------------------------------------------------------------
// Same fields as p2b_challenger (include/plonky2_b200.h) = Challenger {sponge_state, input_buffer, output_buffer}.
struct Challenger {
  u64 state[12];
  u64 in[8];
  u64 out[8];
  u32 in_len, out_len;
};

// duplexing (challenger.rs:130-147): overwrite the first lanes with the buffered inputs, permute, refill the outputs
__device__ __noinline__ void duplex(Challenger* c) {
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = c->state[i];
  const u32 filled = c->in_len;
#pragma unroll
  for (int i = 0; i < 8; i++)  // static indices only: the state must stay in registers
    if ((u32)i < filled) s[i] = c->in[i];
  c->in_len = 0;
  poseidon::permute(s);
#pragma unroll
  for (int i = 0; i < 12; i++) c->state[i] = gl::canon(s[i]);
#pragma unroll
  for (int i = 0; i < 8; i++) c->out[i] = c->state[i];
  c->out_len = 8;
}
__device__ __forceinline__ void observe(Challenger* c, u64 e) {  // challenger.rs:42-51
  c->out_len = 0;
  c->in[c->in_len++] = gl::canon(e);
  if (c->in_len == 8) duplex(c);
}
__device__ __forceinline__ u64 challenge(Challenger* c) {  // challenger.rs:83-93
  if (c->in_len != 0 || c->out_len == 0) duplex(c);
  return c->out[--c->out_len];
}

// One transcript step by a single thread: observe n_obs elements, then squeeze n_out challenges.
//   obs_cs == 0: obs[i];  obs_cs > 0: element i = obs[(i & 1) * obs_cs + (i >> 1)] (extension elements held as two
//   columns, observed as c0, c1 per element: observe_extension_elements, challenger.rs:62-70)
//   modulus != 0: out[i] = challenge % modulus (query indices, prover.rs:185)
__global__ void challenger_step_kernel(Challenger* c, const u64* __restrict__ obs, u32 n_obs, u64 obs_cs,
                                       u64* __restrict__ out, u32 n_out, u64 modulus) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (u32 i = 0; i < n_obs; i++) observe(c, obs_cs ? obs[(u64)(i & 1) * obs_cs + (i >> 1)] : obs[i]);
  for (u32 i = 0; i < n_out; i++) {
    u64 v = challenge(c);
    out[i] = modulus ? v % modulus : v;
  }
}

// Grinding (prover.rs:145-161): candidate w is accepted when perm(state with w at the next input lane)[RATE-1] has
// at least min_lz leading zero bits.  The smallest accepted candidate of the launch is kept (atomicMin), which makes
// the witness deterministic; the reference's rayon find_any may return any accepted candidate.
__global__ void __launch_bounds__(128) pow_search_kernel(const Challenger* __restrict__ c, u64 base, u64 count, u32 min_lz,
                                                         unsigned long long* __restrict__ found) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 cand = base + i;
  if (cand >= gl::P) return;
  u64 s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = c->state[k];
  const u32 pos = c->in_len;
#pragma unroll
  for (int k = 0; k < 8; k++) {  // static indices only: the state must stay in registers
    if ((u32)k < pos) s[k] = c->in[k];
    if ((u32)k == pos) s[k] = cand;
  }
  poseidon::permute(s);
  u64 resp = gl::canon(s[7]);
  u32 lz = resp ? (u32)__clzll((long long)resp) : 64u;
  if (lz >= min_lz) atomicMin(found, (unsigned long long)cand);
}

// out[i] = base^exps[i] in the extension (alpha-power tables and the per-batch weights alpha^count)
__global__ void ext_pows_kernel(const u64* __restrict__ base2, const u64* __restrict__ exps, E2* __restrict__ out, u64 count) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[i] = ecanon(epow(E2{base2[0], base2[1]}, exps[i]));
}

// ---- composition polynomial: comp[i] = sum_j alpha^j f_j[i]  (reduce_polys_base, reducing.rs:87-100) ----
// cols[j]: device pointer to the j-th polynomial's coefficients; out: [2][n]
__global__ void __launch_bounds__(256) reduce_polys_kernel(const u64* const* __restrict__ cols, const E2* __restrict__ apow,
                                                           u32 npolys, u64 n, u64* __restrict__ out) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 a0 = 0, a1 = 0;
  for (u32 j = 0; j < npolys; j++) {
    const u64 x = __ldg(cols[j] + i);
    const E2 a = apow[j];
    a0 = gl::mul_add(a.a, x, a0);
    a1 = gl::mul_add(a.b, x, a1);
  }
  out[i] = gl::canon(a0);
  out[n + i] = gl::canon(a1);
}

// ---- divide_by_linear as a suffix scan (division.rs:75-88) ---------------------------------------------
// S_j = sum_{k >= j} c_k z^(k-j)  (S_j = c_j + z S_{j+1});  quotient q_{j-1} = S_j for j >= 1, and the final polynomial
// is X * quotient (oracle.rs:1084), i.e. final[j] = S_j for j >= 1, final[0] = 0.
// Three phases over chunks of `ch` coefficients: chunk totals, carries across chunks (one CTA), apply.
__global__ void scan_chunk_totals_kernel(const u64* __restrict__ comp, u64 n, u32 ch, E2 z, E2* __restrict__ totals, u64 T) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  u64 lo = t * ch, hi = min(n, lo + ch);
  E2 acc{0, 0};
  for (u64 k = hi; k-- > lo;) acc = eadd(emul(acc, z), E2{comp[k], comp[n + k]});
  totals[t] = acc;
}

// totals[t] = H_t on entry; on exit totals[t] = C_t = S_{(t+1)*ch} = sum_{t' > t} H_t' zc^(t'-t-1), zc = z^ch.
// One CTA of 1024 threads; thread q owns G consecutive chunks.
__global__ void __launch_bounds__(1024) scan_carries_kernel(E2* __restrict__ totals, u64 T, E2 zc) {
  __shared__ E2 sh[1024];
  const u32 q = threadIdx.x;
  const u64 G = (T + 1023) / 1024;
  const u64 t0 = min(T, (u64)q * G), t1 = min(T, t0 + G);
  E2 acc{0, 0};
  for (u64 t = t1; t-- > t0;) acc = eadd(emul(acc, zc), totals[t]);  // L_q = sum H_t zc^(t - t0)
  sh[q] = acc;
  E2 m = epow(zc, G);
  __syncthreads();
  // inclusive suffix scan I_q = sum_{q' >= q} L_q' m^(q'-q)  (Hillis-Steele, multiplier squared every step)
  for (u32 d = 1; d < 1024; d <<= 1) {
    E2 v = sh[q];
    if (q + d < 1024) v = eadd(v, emul(m, sh[q + d]));
    __syncthreads();
    sh[q] = v;
    m = emul(m, m);
    __syncthreads();
  }
  E2 cur = q + 1 < 1024 ? sh[q + 1] : E2{0, 0};  // carry into this thread's top chunk
  for (u64 t = t1; t-- > t0;) {
    E2 h = totals[t];
    totals[t] = cur;
    cur = eadd(h, emul(zc, cur));
  }
}

// out[j] (+)= weight * S_j for the chunk's j (out[0] = 0 contribution), out: [2][n]
__global__ void scan_apply_kernel(const u64* __restrict__ comp, u64 n, u32 ch, E2 z, const E2* __restrict__ carries, u64 T,
                                  const E2* __restrict__ weight, int accumulate, u64* __restrict__ out) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  u64 lo = t * ch, hi = min(n, lo + ch);
  E2 cur = carries[t];
  const E2 w = *weight;
  for (u64 k = hi; k-- > lo;) {
    cur = eadd(emul(cur, z), E2{comp[k], comp[n + k]});
    E2 v = k ? emul(w, cur) : E2{0, 0};
    if (accumulate) v = eadd(v, E2{out[k], out[n + k]});
    out[k] = gl::canon(v.a);
    out[n + k] = gl::canon(v.b);
  }
}

// ---- commit-phase fold: out[i] = sum_j beta^j in[i*arity + j]  (reduce_with_powers, prover.rs:101-108) ----
__global__ void __launch_bounds__(256) fold_kernel(const u64* __restrict__ in, u64 len, u32 arity_bits, const u64* __restrict__ beta2,
                                                   u64* __restrict__ out) {
  const u64 out_len = len >> arity_bits;
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_len) return;
  const E2 beta{beta2[0], beta2[1]};
  const u64 base = i << arity_bits;
  E2 acc{0, 0};
  for (u64 j = (u64)1 << arity_bits; j-- > 0;) acc = eadd(emul(acc, beta), E2{in[base + j], in[len + base + j]});
  out[i] = gl::canon(acc.a);
  out[out_len + i] = gl::canon(acc.b);
}

// ---- OpeningSet: f(z) for every polynomial of a batch (proof.rs:313-319) --------------------------------
// grid (nblk, P), 256 threads.  Block b of polynomial p covers coefficients [b*seg, (b+1)*seg); thread t takes the
// indices base + t + m*256 (coalesced), Horner in z^256, then scales by z^(base + t); CTA-wide sum -> partial[p][b].
__global__ void __launch_bounds__(256) eval_partial_kernel(const u64* __restrict__ coeffs, u64 n, u64 seg, E2 z, E2 z_bd,
                                                           E2* __restrict__ partial) {
  __shared__ E2 sh[256];
  const u32 t = threadIdx.x;
  const u64 p = blockIdx.y, base = (u64)blockIdx.x * seg;
  const u64 end = min(n, base + seg);
  const u64* c = coeffs + p * n;
  E2 acc{0, 0};
  if (base + t < end) {
    u64 last = base + t + ((end - 1 - (base + t)) / 256) * 256;
    for (u64 k = last;; k -= 256) {
      acc = emul_add_base(acc, z_bd, __ldg(c + k));
      if (k < base + t + 256) break;
    }
    acc = emul(acc, epow(z, base + t));
  }
  sh[t] = acc;
  __syncthreads();
  for (u32 d = 128; d > 0; d >>= 1) {
    if (t < d) sh[t] = eadd(sh[t], sh[t + d]);
    __syncthreads();
  }
  if (t == 0) partial[p * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void eval_finish_kernel(const E2* __restrict__ partial, u32 nblk, u64 P, u64* __restrict__ out) {
  u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  E2 acc{0, 0};
  for (u32 b = 0; b < nblk; b++) acc = eadd(acc, partial[p * nblk + b]);
  out[2 * p] = gl::canon(acc.a);
  out[2 * p + 1] = gl::canon(acc.b);
}

}  // namespace fri
