/* One host process driving n devices through the C ABI alone (no Python, no torch, no NCCL): the call a Rust prover makes
 * from from_values_with_gpu (plonky2/src/fri/oracle.rs:279).  Compares the multi-device commit with the single-device one
 * (p2b_commit_from_values on device 0): cap, all leaves, 32 opened rows + Merkle paths.
 *   gcc -O2 -I include tests/c/mgpu_commit_test.c -o t -L plonky2-gpu_b200 -lplonky2_b200 -Wl,-rpath,$PWD/plonky2-gpu_b200
 *   ./t <n_devices> [degree_log=12] [polys=135]                                                       */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "plonky2_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != P2B_OK) { fprintf(stderr, "%s failed: %d %s\n", #x, rc_, p2b_last_error()); return 1; } } while (0)

static uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  int ndev = argc > 1 ? atoi(argv[1]) : 2;
  uint32_t k = argc > 2 ? (uint32_t)atoi(argv[2]) : 12, rate_bits = 3, cap_height = 4;
  uint64_t P = argc > 3 ? (uint64_t)atoll(argv[3]) : 135, n = 1ull << k, N = n << rate_bits;
  uint64_t *values, *coeffs_m, *coeffs_s;
  CHECK(p2b_malloc_host(P * n * 8, (void**)&values));
  CHECK(p2b_malloc_host(P * n * 8, (void**)&coeffs_m));
  coeffs_s = malloc(P * n * 8);
  uint64_t seed = 42;
  for (uint64_t i = 0; i < P * n; i++) {
    uint64_t v;
    do v = splitmix(&seed); while (v >= 0xFFFFFFFF00000001ull);
    values[i] = v;
  }
  /* single device */
  p2b_ctx* ctx;
  p2b_batch* single;
  CHECK(p2b_ctx_create(0, &ctx));
  CHECK(p2b_commit_from_values(ctx, values, 1, k, P, rate_bits, cap_height, NULL, 0, &single));
  uint64_t cap_s[16 * 4], cap_m[16 * 4];
  CHECK(p2b_batch_get_cap(single, cap_s));
  CHECK(p2b_batch_get_coeffs(single, coeffs_s));
  /* n devices, one process */
  p2b_mgpu* g;
  p2b_mgpu_batch* multi;
  CHECK(p2b_mgpu_create(NULL, ndev, &g));
  for (int rep = 0; rep < 3; rep++) { /* repeated commits reuse events and buffers */
    CHECK(p2b_mgpu_commit_from_values(g, values, k, P, rate_bits, cap_height, coeffs_m, &multi));
    CHECK(p2b_mgpu_batch_get_cap(multi, cap_m));
    if (rep < 2) p2b_mgpu_batch_destroy(multi);
  }
  int bad = memcmp(cap_s, cap_m, sizeof(cap_s)) != 0;
  if (bad) fprintf(stderr, "cap mismatch\n");
  if (memcmp(coeffs_s, coeffs_m, P * n * 8)) { fprintf(stderr, "coefficients mismatch\n"); bad = 1; }
  uint64_t* rows_s = malloc(N * P * 8), *rows_m = malloc(N * P * 8);
  CHECK(p2b_batch_get_leaves(single, 0, N, rows_s));
  CHECK(p2b_mgpu_batch_get_leaves(multi, 0, N, rows_m));
  if (memcmp(rows_s, rows_m, N * P * 8)) { fprintf(stderr, "leaves mismatch\n"); bad = 1; }
  uint64_t idx[32], layers = k + rate_bits - cap_height;
  for (int i = 0; i < 32; i++) idx[i] = splitmix(&seed) % N;
  uint64_t *or_s = malloc(32 * P * 8), *or_m = malloc(32 * P * 8), *sb_s = malloc(32 * layers * 32), *sb_m = malloc(32 * layers * 32);
  CHECK(p2b_batch_open_rows(single, idx, 32, or_s, sb_s));
  CHECK(p2b_mgpu_batch_open_rows(multi, idx, 32, or_m, sb_m));
  if (memcmp(or_s, or_m, 32 * P * 8) || memcmp(sb_s, sb_m, 32 * layers * 32)) { fprintf(stderr, "opened rows / paths mismatch\n"); bad = 1; }
  printf("%s: %d device(s), peer access %d, 2^%u x %llu, cap word0 %016llx\n", bad ? "FAIL" : "OK", p2b_mgpu_device_count(g),
         p2b_mgpu_peer_access(g), k, (unsigned long long)P, (unsigned long long)cap_m[0]);
  p2b_mgpu_batch_destroy(multi);
  p2b_mgpu_destroy(g);
  p2b_batch_destroy(single);
  p2b_ctx_destroy(ctx);
  return bad;
}
