// Host check of csrc/mds_fft.cuh against the direct circulant form (poseidon.rs:172-260) for T = int64_t, uint32_t
// (wrapping) and double (exactness bound of the FP64-pipe path): prints "ok" and exits 0.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../plonky2-gpu_b200/csrc/mds_fft.cuh"
static const int64_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
// magnitude bound: a number type whose every operation adds absolute values, so evaluating mds12 on |x_i| <= B yields an
// upper bound of every intermediate of the real computation (needed: < 2^53 for the FP64 path to be exact)
struct Abs {
  double v;
  Abs() : v(0) {}
  explicit Abs(double x) : v(x < 0 ? -x : x) {}
  explicit Abs(int k) : v(k < 0 ? -k : k) {}
};
static double g_max = 0;
static Abs note(Abs a) { if (a.v > g_max) g_max = a.v; return a; }
static Abs operator+(Abs a, Abs b) { return note(Abs(a.v + b.v)); }
static Abs operator-(Abs a, Abs b) { return note(Abs(a.v + b.v)); }
static Abs operator-(Abs a) { return a; }
static Abs operator*(Abs a, Abs b) { return note(Abs(a.v * b.v)); }

int main() {
  {
    Abs x[12], y[12];
    for (int i = 0; i < 12; i++) x[i] = Abs(8589934592.0);  // 2^33: a 32-bit half plus a folded 32-bit constant half
    mdsfft::mds12<Abs>(x, y);
    printf("max |intermediate| for |x| <= 2^33: %.0f = 2^%.2f (must be < 2^53)\n", g_max, __builtin_log2(g_max));
    if (!(g_max < 9007199254740992.0)) return 1;
  }
  // every vertex of the box [0, 2^32 - 1]^12 (extremes of all linear forms)
  for (int mask = 0; mask < 4096; mask++) {
    int64_t x[12], got[12];
    double xd[12], yd[12];
    for (int i = 0; i < 12; i++) { x[i] = (mask >> i) & 1 ? 4294967295ll : 0; xd[i] = (double)x[i]; }
    mdsfft::mds12<int64_t>(x, got);
    mdsfft::mds12<double>(xd, yd);
    for (int r = 0; r < 12; r++) {
      int64_t s = 0;
      for (int i = 0; i < 12; i++) s += C[i] * x[(i + r) % 12];
      if (r == 0) s += 8 * x[0];
      if (got[r] != s || yd[r] != (double)s) { printf("vertex mismatch mask %d r %d\n", mask, r); return 1; }
    }
  }
  for (int it = 0; it < 200000; it++) {
    int64_t x[12], want[12], got[12];
    int bits = 1 + it % 34;  // up to 2^34: the halves plus a folded round-constant half
    for (int i = 0; i < 12; i++) {
      x[i] = (int64_t)(rnd() >> (64 - bits));
      if (it % 7 == 0) x[i] = ((int64_t)1 << bits) - 1;  // extreme
    }
    for (int r = 0; r < 12; r++) {
      int64_t s = 0;
      for (int i = 0; i < 12; i++) s += C[i] * x[(i + r) % 12];
      if (r == 0) s += 8 * x[0];
      want[r] = s;
    }
    mdsfft::mds12<int64_t>(x, got);
    for (int r = 0; r < 12; r++) if (got[r] != want[r]) { printf("int64 mismatch it %d r %d\n", it, r); return 1; }
    double xd[12], yd[12];
    for (int i = 0; i < 12; i++) xd[i] = (double)x[i];
    mdsfft::mds12<double>(xd, yd);
    for (int r = 0; r < 12; r++) if (yd[r] != (double)want[r] || (int64_t)yd[r] != want[r]) { printf("double mismatch it %d r %d\n", it, r); return 1; }
    if (bits <= 22) {
      uint32_t xu[12], yu[12];
      for (int i = 0; i < 12; i++) xu[i] = (uint32_t)x[i];
      mdsfft::mds12<uint32_t>(xu, yu);
      for (int r = 0; r < 12; r++) if (yu[r] != (uint32_t)want[r]) { printf("u32 mismatch it %d r %d\n", it, r); return 1; }
    }
  }
  printf("ok\n");
  return 0;
}
