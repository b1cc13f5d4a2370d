// plonky2_b200.cu -- the library's single translation unit: context, pass planner, commit orchestration
// and the C ABI declared in include/plonky2_b200.h.
//
// Path implemented (reference file:line):
//   PolynomialBatch::from_values   plonky2/src/fri/oracle.rs:709-731
//   PolynomialBatch::from_coeffs   plonky2/src/fri/oracle.rs:911-977  (lde_values :979-1004)
//   MerkleTree::new / prove        plonky2/src/hash/merkle_tree.rs:283-319, 392-440
//   get_lde_values                 plonky2/src/fri/oracle.rs:1007-1018
// and the boundary the reference's Rust side binds: cuda/src/lib.rs:52-145.
//
// HBM layout of a committed batch (one cudaMallocAsync block each, 256-byte aligned):
//   coeffs  [P][n]            column-major, natural coefficient order           8 n P bytes
//   leaves  [N][P + salt]     row-major, reference leaf order (bit-reversed)     8 N (P+salt) bytes
//   digests [2 (N - 2^c)][4]  reference recursive layout                         64 (N - 2^c) bytes
//   cap     [2^c][4]
// plus a context-owned scratch [P][n] that holds one coset block between its NTT passes.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX3: ranges cost nothing unless a profiler injects its library

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <atomic>
#include <mutex>
#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/plonky2_b200.h"
#include "merkle.cuh"
#include "ntt.cuh"
#include "quotient.cuh"
#include "fri.cuh"
#include "permutation.cuh"

using gl::u32;
using gl::u64;

// ======================================================================================================
// errors
// ======================================================================================================
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
// a failed allocation reports what the device had left (the figure a caller needs to size its batches)
static int fail_cuda(cudaError_t e, const char* expr, const char* file, int line) {
  if (e != cudaErrorMemoryAllocation) return fail(P2B_ERR_CUDA, "%s failed: %s (%s:%d)", expr, cudaGetErrorString(e), file, line);
  int dev = -1;
  size_t fr = 0, tot = 0;
  cudaGetLastError();
  cudaGetDevice(&dev);
  cudaMemGetInfo(&fr, &tot);
  return fail(P2B_ERR_OOM, "%s failed: %s; device %d reports %llu of %llu MiB free (%s:%d)", expr, cudaGetErrorString(e), dev,
              (unsigned long long)(fr >> 20), (unsigned long long)(tot >> 20), file, line);
}
#define CUDA_TRY(expr)                                                  \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)
// Stream-ordered allocation that survives the allocator's transient failures.  cudaMallocAsync can report "out of memory" on
// an almost empty device: observed on 2 x B200 with peer-mapped default pools (cudaMemPoolSetAccess, mgpu.cuh) -- the pool held
// 5984 MiB reserved / 5591 MiB used, 175 GB of the device were free, and growing the pool by 1.1 GB failed, also after a device
// synchronisation, until the unused reserve was handed back (cudaMemPoolTrimTo).  So: retry after draining the device (blocks
// freed on other streams become reusable), then once more after trimming the pool; a third failure is the real thing.
static std::atomic<unsigned long long> g_pool_retries{0};
static cudaError_t pool_alloc(void** out, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMallocAsync(out, bytes, st);
  if (e != cudaErrorMemoryAllocation) return e;
  cudaGetLastError();
  cudaDeviceSynchronize();
  g_pool_retries++;
  e = cudaMallocAsync(out, bytes, st);
  if (e != cudaErrorMemoryAllocation) return e;
  cudaGetLastError();
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    cudaMemPoolTrimTo(pool, 0);
    g_pool_retries++;
    e = cudaMallocAsync(out, bytes, st);
  }
  return e;
}
template <class T>
static cudaError_t pool_alloc(T** out, size_t bytes, cudaStream_t st) { return pool_alloc(reinterpret_cast<void**>(out), bytes, st); }
extern "C" unsigned long long p2b_debug_pool_retries(void) { return g_pool_retries.load(); }
#define P2B_TRY(expr)           \
  do {                          \
    int _rc = (expr);           \
    if (_rc != P2B_OK) return _rc; \
  } while (0)

extern "C" const char* p2b_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* p2b_version(void) { return "plonky2_b200 0.1 sm_100a"; }

// ======================================================================================================
// host-side field helpers (table set-up only)
// ======================================================================================================
namespace hostf {
typedef unsigned __int128 u128;
static inline u64 mul(u64 a, u64 b) { return (u64)(((u128)a * b) % gl::P); }
static inline u64 pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mul(r, b);
    b = mul(b, b);
    e >>= 1;
  }
  return r;
}
static inline u64 inv(u64 a) { return pow(a, gl::P - 2); }
// primitive_root_of_unity (field/src/types.rs:268-272), POWER_OF_TWO_GENERATOR goldilocks_field.rs:89
static inline u64 root(unsigned n_log) {
  u64 b = 1753635133440165772ull;
  for (unsigned i = 0; i < 32 - n_log; i++) b = mul(b, b);
  return b;
}
static constexpr u64 COSET_SHIFT = 7;  // MULTIPLICATIVE_GROUP_GENERATOR, goldilocks_field.rs:82
}  // namespace hostf

// ======================================================================================================
// context
// ======================================================================================================
struct p2b_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // NTT passes, tree layers, copies
  cudaStream_t stream2 = nullptr;  // leaf hashing of block b overlaps the NTT of block b+1
  cudaStream_t stream_h2d = nullptr, stream_d2h = nullptr;  // host copies overlapped with compute (one DMA engine each way)
  cudaEvent_t ev_copy[32] = {};
  bool owns_streams = true;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  // twiddle tables U[b], b < 2^tw_log  (forward and inverse roots)
  u64* U_fwd = nullptr;
  u64* U_inv = nullptr;
  u64* d_roots = nullptr;  // [2][34]
  u32 tw_log = 0;
  // scratch
  u64* scratch = nullptr;
  u64 scratch_elems = 0;
  u64 launches = 0;
  bool debug_force_redo = false;  // test hook (p2b_ctx_debug_force_exact_redo)
  // optional per-kernel timing of the dominant kernel (leaf hashing): event pairs on the launching stream
  bool time_hash = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> hash_events;
  size_t hash_events_used = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  // stage tracing (p2b_ctx_trace): the reference's TimingTree scopes (plonky2/src/util/timing.rs:8-192; "IFFT",
  // "FFT + blinding", "build Merkle tree" at fri/oracle.rs:717-966, ...) as CUDA-event pairs on the launching stream
  bool trace = false;
  struct TraceSpan { const char* name; cudaEvent_t a, b; };
  std::vector<TraceSpan> spans;
  size_t spans_used = 0;
};

// RAII stage marker: an NVTX range around the host-side enqueue (shows up in nsys / ncu timelines) and, when tracing is
// enabled on the context, an event pair on `st` whose elapsed time p2b_ctx_trace_report() sums per stage name.
struct Stage {
  p2b_ctx* c;
  cudaStream_t st;
  p2b_ctx::TraceSpan* span = nullptr;
  Stage(p2b_ctx* c_, cudaStream_t st_, const char* name) : c(c_), st(st_) {
    nvtxRangePushA(name);
    if (!c->trace) return;
    if (c->spans_used == c->spans.size()) {
      p2b_ctx::TraceSpan sp{name, nullptr, nullptr};
      if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
      c->spans.push_back(sp);
    }
    span = &c->spans[c->spans_used++];
    span->name = name;
    cudaEventRecord(span->a, st);
  }
  ~Stage() {
    if (span) cudaEventRecord(span->b, st);
    nvtxRangePop();
  }
};

// Live contexts: objects that outlive their context (a binding's finalizers run in any order at process exit) must not touch
// its streams -- their destructors check here and only drop the host-side record then.
static std::mutex g_live_mu;
static std::vector<p2b_ctx*> g_live_ctx;
static void ctx_register(p2b_ctx* c) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  g_live_ctx.push_back(c);
}
static void ctx_unregister(p2b_ctx* c) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  g_live_ctx.erase(std::remove(g_live_ctx.begin(), g_live_ctx.end(), c), g_live_ctx.end());
}
static bool ctx_alive(const p2b_ctx* c) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  return std::find(g_live_ctx.begin(), g_live_ctx.end(), c) != g_live_ctx.end();
}

static int ensure_scratch(p2b_ctx* c, u64 elems) {
  if (c->scratch_elems >= elems) return P2B_OK;
  if (c->scratch) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream2));
    CUDA_TRY(cudaFree(c->scratch));
    c->scratch = nullptr;
    c->scratch_elems = 0;
  }
  CUDA_TRY(cudaMalloc(&c->scratch, elems * sizeof(u64)));
  c->scratch_elems = elems;
  return P2B_OK;
}

static int ensure_twiddles(p2b_ctx* c, u32 log_count) {
  if (log_count < 4) log_count = 4;
  if (c->tw_log >= log_count) return P2B_OK;
  if (log_count > 31) return fail(P2B_ERR_INVALID, "transform too large: twiddle table 2^%u", log_count);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream2));
  if (c->U_fwd) CUDA_TRY(cudaFree(c->U_fwd));
  if (c->U_inv) CUDA_TRY(cudaFree(c->U_inv));
  c->U_fwd = c->U_inv = nullptr;
  c->tw_log = 0;
  u64 count = (u64)1 << log_count;
  CUDA_TRY(cudaMalloc(&c->U_fwd, count * sizeof(u64)));
  CUDA_TRY(cudaMalloc(&c->U_inv, count * sizeof(u64)));
  if (!c->d_roots) {
    u64 h[2][34];
    for (int j = 0; j < 34; j++) {
      u64 r = j <= 32 ? hostf::root(j) : 1;
      h[0][j] = r;
      h[1][j] = hostf::inv(r);
    }
    CUDA_TRY(cudaMalloc(&c->d_roots, sizeof(h)));
    CUDA_TRY(cudaMemcpyAsync(c->d_roots, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  unsigned blocks = (unsigned)((count + 255) / 256);
  ntt::build_twiddles_kernel<<<blocks, 256, 0, c->stream>>>(c->U_fwd, count, c->d_roots);
  ntt::build_twiddles_kernel<<<blocks, 256, 0, c->stream>>>(c->U_inv, count, c->d_roots + 34);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  c->tw_log = log_count;
  return P2B_OK;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return P2B_OK;
}

static int ctx_init_common(p2b_ctx* c) {
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10)
    return fail(P2B_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", c->device,
                prop.major, prop.minor);
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreate(&c->ev_t0));
  CUDA_TRY(cudaEventCreate(&c->ev_t1));
  for (auto& e : c->ev_copy) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_TRY(poseidon::upload_constants());
  // let the stream-ordered pool keep its memory between commits (no OS round trip inside a timed step)
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, c->device));
  uint64_t thresh = UINT64_MAX;
  CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  return P2B_OK;
}

extern "C" int p2b_ctx_create(int device, p2b_ctx** out) {
  if (!out) return fail(P2B_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(P2B_ERR_CUDA, "no CUDA device");
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= ndev) return fail(P2B_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  p2b_ctx* c = new (std::nothrow) p2b_ctx();
  if (!c) return fail(P2B_ERR_OOM, "host allocation failed");
  c->device = device;
  int rc = P2B_OK;
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking)) != cudaSuccess)
    rc = fail(P2B_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  if (rc == P2B_OK) rc = ctx_init_common(c);
  if (rc != P2B_OK) {
    delete c;
    return rc;
  }
  ctx_register(c);
  *out = c;
  return P2B_OK;
}

extern "C" void p2b_ctx_destroy(p2b_ctx* c) {
  if (!c || !ctx_alive(c)) return;
  ctx_unregister(c);
  cudaSetDevice(c->device);
  for (cudaStream_t st : {c->stream, c->stream2, c->stream_h2d, c->stream_d2h})   // all four: copies may still be in flight
    if (st) cudaStreamSynchronize(st);
  cudaFree(c->U_fwd);
  cudaFree(c->U_inv);
  cudaFree(c->d_roots);
  cudaFree(c->scratch);
  for (cudaEvent_t e : {c->ev_a, c->ev_b, c->ev_t0, c->ev_t1})
    if (e) cudaEventDestroy(e);
  for (auto& pr : c->hash_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (cudaEvent_t e : c->ev_copy)
    if (e) cudaEventDestroy(e);
  for (auto& sp : c->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  if (c->stream_h2d) cudaStreamDestroy(c->stream_h2d);
  if (c->stream_d2h) cudaStreamDestroy(c->stream_d2h);
  if (c->owns_streams) {
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
  }
  delete c;
}

extern "C" void* p2b_ctx_stream(p2b_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int p2b_ctx_synchronize(p2b_ctx* c) {
  if (!c) return fail(P2B_ERR_INVALID, "ctx is NULL");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream2));
  CUDA_TRY(cudaStreamSynchronize(c->stream_h2d));
  CUDA_TRY(cudaStreamSynchronize(c->stream_d2h));
  return P2B_OK;
}
extern "C" uint64_t p2b_ctx_launch_count(const p2b_ctx* c) { return c ? c->launches : 0; }

extern "C" int p2b_ctx_trace(p2b_ctx* c, int enable) {
  if (!c) return fail(P2B_ERR_INVALID, "NULL context");
  c->trace = enable != 0;
  c->spans_used = 0;
  return P2B_OK;
}
// Synchronises the context and writes one line per stage, "name<TAB>calls<TAB>total ms", in first-seen order.  Stages on
// different streams overlap (leaf hashing runs beside the next block's NTT), so the totals can exceed the wall time.
extern "C" int p2b_ctx_trace_report(p2b_ctx* c, char* buf, uint64_t buf_len) {
  if (!c || !buf || buf_len == 0) return fail(P2B_ERR_INVALID, "NULL argument");
  P2B_TRY(p2b_ctx_synchronize(c));
  std::vector<const char*> names;
  std::vector<double> ms_sum;
  std::vector<u64> calls;
  for (size_t i = 0; i < c->spans_used; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->spans[i].a, c->spans[i].b) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    size_t k = 0;
    while (k < names.size() && strcmp(names[k], c->spans[i].name)) k++;
    if (k == names.size()) {
      names.push_back(c->spans[i].name);
      ms_sum.push_back(0);
      calls.push_back(0);
    }
    ms_sum[k] += ms;
    calls[k]++;
  }
  std::string out;
  char line[256];
  for (size_t k = 0; k < names.size(); k++) {
    snprintf(line, sizeof(line), "%s\t%llu\t%.4f\n", names[k], (unsigned long long)calls[k], ms_sum[k]);
    out += line;
  }
  if (out.size() + 1 > buf_len) return fail(P2B_ERR_INVALID, "trace report needs %zu bytes", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  c->spans_used = 0;
  return P2B_OK;
}
extern "C" int p2b_ctx_debug_force_exact_redo(p2b_ctx* c, int enable) {
  if (!c) return fail(P2B_ERR_INVALID, "ctx is NULL");
  c->debug_force_redo = enable != 0;
  return P2B_OK;
}
extern "C" int p2b_ctx_time_leaf_hash(p2b_ctx* c, int enable) {
  if (!c) return fail(P2B_ERR_INVALID, "ctx is NULL");
  c->time_hash = enable != 0;
  c->hash_events_used = 0;
  return P2B_OK;
}
extern "C" int p2b_ctx_leaf_hash_time(p2b_ctx* c, double* total_ms, uint64_t* launches) {
  if (!c || !total_ms || !launches) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream2));
  double t = 0;
  for (size_t i = 0; i < c->hash_events_used; i++) {
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->hash_events[i].first, c->hash_events[i].second));
    t += ms;
  }
  *total_ms = t;
  *launches = c->hash_events_used;
  return P2B_OK;
}

// ======================================================================================================
// pass planner
// ======================================================================================================
#ifndef P2B_FINAL_L
#define P2B_FINAL_L 10
#endif
#ifndef P2B_STRIDED_T16_MAX_SMEM
#define P2B_STRIDED_T16_MAX_SMEM (100 * 1024)  // above this a T = 16 tile leaves one CTA per SM: use T = 8
#endif
static constexpr u32 FINAL_L = P2B_FINAL_L;  // levels of the last pass (2^FINAL_L-row chunks)
static constexpr u32 STRIDED_MAX_L = 11;
static constexpr int FINAL_C = 8;      // columns per CTA in the final LDE / column-major pass
static constexpr int INTT_J = 16;      // chunks per CTA in the inverse final pass
static constexpr u64 HASH_LAUNCH_MIN_LEAVES = (u64)1 << 20;   // ~8 waves of 148 x 7 CTAs x 128 leaves

struct Plan {
  u32 k;
  u32 n_strided;
  u32 s0[4], L[4];  // strided passes
  u32 final_s0, final_L;
};
static Plan make_plan(u32 k) {
  Plan p{};
  p.k = k;
  p.final_L = k < FINAL_L ? k : FINAL_L;
  p.final_s0 = k - p.final_L;
  u32 rem = p.final_s0;
  p.n_strided = (rem + STRIDED_MAX_L - 1) / STRIDED_MAX_L;
  u32 s = 0;
  for (u32 i = 0; i < p.n_strided; i++) {
    u32 left = p.n_strided - i;
    u32 L = (rem + left - 1) / left;
    p.s0[i] = s;
    p.L[i] = L;
    s += L;
    rem -= L;
  }
  return p;
}

static int launch_strided(p2b_ctx* c, cudaStream_t st, const ntt::PassArgs& a, const ntt::LevelScale& sc) {
  const u64 stride = ((u64)1 << a.k) >> (a.s0 + a.L);
  // T = 16 positions (128-byte segments) unless the tile would not fit in shared memory
  size_t smem16 = ((size_t)(16 + 1) << a.L) * 8;
  bool use16 = stride >= 16 && smem16 <= c->smem_optin && smem16 <= (size_t)P2B_STRIDED_T16_MAX_SMEM;
  int T = use16 ? 16 : 8;
  if (stride < (u64)T) return fail(P2B_ERR_INVALID, "internal: strided pass with stride %llu", (unsigned long long)stride);
  size_t smem = ((size_t)(T + 1) << a.L) * 8;
  if (smem > c->smem_optin) return fail(P2B_ERR_INVALID, "internal: strided tile does not fit shared memory");
  u64 tiles = ((u64)1 << a.s0) * (stride / T);
  dim3 grid((unsigned)tiles, a.ncols);
  if (use16) {
    P2B_TRY(opt_in_smem(ntt::ntt_strided_pass_kernel<16>, smem));
    ntt::ntt_strided_pass_kernel<16><<<grid, 512, smem, st>>>(a, sc);
  } else {
    P2B_TRY(opt_in_smem(ntt::ntt_strided_pass_kernel<8>, smem));
    ntt::ntt_strided_pass_kernel<8><<<grid, 512, smem, st>>>(a, sc);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

// ======================================================================================================
// inverse NTT of P columns: src [P][n] -> dst [P][n] (natural order, scaled).  tmp: [P][n] scratch, needed
// when the plan has strided passes and src must be preserved (tmp may equal dst when src != dst is not
// required to survive).
// ======================================================================================================
static int run_ifft(p2b_ctx* c, const u64* src, u64* dst, u64* tmp, u32 k, u64 P) {
  if (P == 0) return P2B_OK;
  Stage stage(c, c->stream, "IFFT");   // fri/oracle.rs:717-721
  P2B_TRY(ensure_twiddles(c, k > 0 ? k - 1 : 0));
  Plan pl = make_plan(k);
  const u64 n = (u64)1 << k;
  ntt::LevelScale sc{};
  ntt::PassArgs a{};
  a.k = k;
  a.block_base = 0;
  a.U = c->U_inv;
  a.scaled = 0;
  a.ncols = (u32)P;
  a.force_redo = c->debug_force_redo;
  const u64* cur = src;
  for (u32 i = 0; i < pl.n_strided; i++) {
    a.src = cur;
    a.dst = tmp;
    a.src_cs = n;
    a.dst_cs = n;
    a.s0 = pl.s0[i];
    a.L = pl.L[i];
    P2B_TRY(launch_strided(c, c->stream, a, sc));
    cur = tmp;
  }
  a.src = cur;
  a.dst = dst;
  a.src_cs = a.dst_cs = n;
  a.s0 = pl.final_s0;
  a.L = pl.final_L;
  u64 chunks = (u64)1 << a.s0;
  if (a.L <= 9) {
    size_t smem = ((size_t)INTT_J << a.L) * 8;
    P2B_TRY(opt_in_smem(ntt::intt_final_pass_kernel<INTT_J>, smem));
    dim3 grid((unsigned)((chunks + INTT_J - 1) / INTT_J), (unsigned)P);
    ntt::intt_final_pass_kernel<INTT_J><<<grid, 512, smem, c->stream>>>(a, gl::inverse_2exp(k));
  } else {  // larger chunks: 8 lanes (64-byte output segments) keep the tile at 2^L * 64 bytes
    size_t smem = ((size_t)8 << a.L) * 8;
    P2B_TRY(opt_in_smem(ntt::intt_final_pass_kernel<8>, smem));
    dim3 grid((unsigned)((chunks + 7) / 8), (unsigned)P);
    ntt::intt_final_pass_kernel<8><<<grid, 512, smem, c->stream>>>(a, gl::inverse_2exp(k));
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

// sigma[s] = shift^(n / 2^(s+1)) for the LDE sub-problem of size n = 2^k on the coset shift * H
static ntt::LevelScale lde_scale(u32 k, u64 shift = hostf::COSET_SHIFT) {
  ntt::LevelScale sc{};
  for (u32 s = 0; s < k; s++) sc.sigma[s] = hostf::pow(shift, ((u64)1 << k) >> (s + 1));
  return sc;
}

// One coset block b of the LDE: coeffs [P][n] -> rows [b*n, (b+1)*n) of leaves (row-major).
static int run_lde_block(p2b_ctx* c, cudaStream_t st, const u64* coeffs, u64 coeffs_cs, u64* tmp, u32 k, u64 P,
                         u64 b, const ntt::LevelScale& sc, u64* leaves, u64 row_stride, u64 col0, u64 row0) {
  Stage stage(c, st, "FFT + blinding");   // fri/oracle.rs:927-931 (LDE of one coset block, written as leaf rows)
  Plan pl = make_plan(k);
  const u64 n = (u64)1 << k;
  ntt::PassArgs a{};
  a.k = k;
  a.block_base = b;
  a.U = c->U_fwd;
  a.scaled = 1;
  a.ncols = (u32)P;
  a.force_redo = c->debug_force_redo;
  const u64* cur = coeffs;
  u64 cur_cs = coeffs_cs;
  for (u32 i = 0; i < pl.n_strided; i++) {
    a.src = cur;
    a.src_cs = cur_cs;
    a.dst = tmp;
    a.dst_cs = n;
    a.s0 = pl.s0[i];
    a.L = pl.L[i];
    P2B_TRY(launch_strided(c, st, a, sc));
    cur = tmp;
    cur_cs = n;
  }
  a.src = cur;
  a.src_cs = cur_cs;
  a.dst = leaves;
  a.dst_cs = 0;
  a.s0 = pl.final_s0;
  a.L = pl.final_L;
  size_t smem = (((size_t)(FINAL_C + 1) << a.L) + ((size_t)1 << a.L)) * 8;
  P2B_TRY(opt_in_smem(ntt::ntt_final_pass_kernel<FINAL_C, ntt::MODE_ROWS>, smem));
  const u64 tiles = ((u64)1 << a.s0) * ((P + FINAL_C - 1) / FINAL_C);   // column group fastest (see the kernel's header)
  if (tiles > 0x7fffffffull) return fail(P2B_ERR_UNSUPPORTED, "LDE final pass: %llu tiles exceed the grid limit", (unsigned long long)tiles);
  ntt::ntt_final_pass_kernel<FINAL_C, ntt::MODE_ROWS><<<(unsigned)tiles, 512, smem, st>>>(a, sc, row0, row_stride, col0);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

// ======================================================================================================
// Merkle tree over leaf rows
// ======================================================================================================
static int launch_hash_leaves(p2b_ctx* c, cudaStream_t st, const u64* leaves, u64 row_stride, u64 col_stride,
                              u32 leaf_len, u64 first_leaf, u64 count, const merkle::TreeShape& shape, u64* digests,
                              u64* cap) {
  if (count == 0) return P2B_OK;
  Stage stage(c, st, "build Merkle tree: leaf hashes");   // fri/oracle.rs:962-966, merkle_tree.rs:210-244
  unsigned blocks = (unsigned)((count + P2B_HASH_BLOCK - 1) / P2B_HASH_BLOCK);
  // `leaves` points at the first leaf of this launch; first_leaf is its global index in the tree
  std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
  if (c->time_hash) {
    if (c->hash_events_used == c->hash_events.size()) {
      cudaEvent_t a, b;
      CUDA_TRY(cudaEventCreate(&a));
      CUDA_TRY(cudaEventCreate(&b));
      c->hash_events.emplace_back(a, b);
    }
    ev = &c->hash_events[c->hash_events_used++];
    CUDA_TRY(cudaEventRecord(ev->first, st));
  }
  merkle::hash_leaves_kernel<<<blocks, P2B_HASH_BLOCK, 0, st>>>(leaves, row_stride, col_stride, leaf_len, count, first_leaf, shape,
                                                                digests, cap);
  if (ev) CUDA_TRY(cudaEventRecord(ev->second, st));
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}
// Digest layers above the leaves [leaf0, leaf1) for as long as whole nodes fit that range, starting above
// `from_layer`.  Returns the last layer computed (== shape.sub_log when the range reaches the cap).
static int launch_layers(p2b_ctx* c, cudaStream_t st, const merkle::TreeShape& shape, u64* digests, u64* cap,
                         u64 leaf0, u64 leaf1, u32 from_layer, u32* top_layer) {
  Stage stage(c, st, "build Merkle tree: digest layers");
  u32 l = from_layer;
  while (l < shape.sub_log) {
    u64 span = (u64)1 << (l + 1);
    if ((leaf0 % span) || (leaf1 % span) || leaf1 <= leaf0) break;
    l++;
    u64 node0 = leaf0 >> l, count = (leaf1 - leaf0) >> l;
    unsigned blocks = (unsigned)((count + P2B_HASH_BLOCK - 1) / P2B_HASH_BLOCK);
    merkle::merkle_layer_kernel<<<blocks, P2B_HASH_BLOCK, 0, st>>>(shape, l, node0, count, digests, cap);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  if (top_layer) *top_layer = l;
  return P2B_OK;
}

// leaf row L, column P + c  <-  salt[c][reverse_bits(L)]   (oracle.rs:998-1002 then :942-952)
__global__ void scatter_salt_kernel(const u64* __restrict__ salt, u64 N, u32 log_N, u64 leaf0, u64 count,
                                    u64* __restrict__ leaves, u64 row_stride, u64 col0) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 L = leaf0 + i;
  u64 src = log_N ? (__brevll(L) >> (64 - log_N)) : 0;
#pragma unroll
  for (int cidx = 0; cidx < P2B_SALT_SIZE; cidx++) leaves[i * row_stride + col0 + cidx] = gl::canon(salt[(u64)cidx * N + src]);
}

// leaf row leaf0 + i, column col0 + c  <-  salt[c][leaf0 + i]   (the reference's device layout, compat.cuh)
__global__ void gather_salt_by_leaf_kernel(const u64* __restrict__ salt, u64 N, u64 leaf0, u64 count, u32 ncols,
                                           u64* __restrict__ rows, u64 row_stride, u64 col0) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  for (u32 cidx = 0; cidx < ncols; cidx++) rows[i * row_stride + col0 + cidx] = gl::canon(salt[(u64)cidx * N + leaf0 + i]);
}

// ======================================================================================================
// batch
// ======================================================================================================
struct p2b_batch {
  p2b_ctx* ctx = nullptr;
  p2b_batch_info info{};
  u64 first_leaf = 0, local_leaves = 0;  // rows held by this (possibly sharded) batch: [first_leaf, first_leaf + local_leaves)
  u32 top_layer = 0;                     // highest digest layer computed from local leaves
  u64* coeffs = nullptr;
  u64* leaves = nullptr;
  u64* digests = nullptr;
  u64* cap = nullptr;
  merkle::TreeShape shape{};
  // pipelined commit (p2b_commit_blocks_begin / _absorb / _finish): sponge states of the local leaves, columns absorbed
  u64* sponge_state = nullptr;
  u64 cols_done = 0;
  bool pipelined = false;
};

static int batch_free(p2b_batch* b) {
  if (!b) return P2B_OK;
  if (!ctx_alive(b->ctx)) {   // the context went first: its device memory is gone with the process / pool, only the record is left
    delete b;
    return P2B_OK;
  }
  cudaSetDevice(b->ctx->device);
  cudaStream_t st = b->ctx->stream;
  if (b->coeffs) cudaFreeAsync(b->coeffs, st);
  if (b->leaves) cudaFreeAsync(b->leaves, st);
  if (b->digests) cudaFreeAsync(b->digests, st);
  if (b->cap) cudaFreeAsync(b->cap, st);
  if (b->sponge_state) cudaFreeAsync(b->sponge_state, st);
  delete b;
  return P2B_OK;
}
extern "C" void p2b_batch_destroy(p2b_batch* b) { batch_free(b); }

// The LDE + Merkle part shared by from_values / from_coeffs and the compat shims.
//   coeffs_d [P][n] (device) -> leaves_d, digests_d, cap_d.   `descending`: process coset blocks b = R-1..0
//   (needed when leaves_d aliases coeffs_d: block 0's rows overwrite the coefficients last).
static int lde_and_merkle(p2b_ctx* c, const u64* coeffs_d, u32 k, u64 P, u32 rate_bits, u32 cap_height,
                          const u64* salt_d, u64* tmp, u64* leaves_d, u64 leaf_len, u64* digests_d, u64* cap_d,
                          bool descending, cudaStream_t wait_before_block0, u64 block_first, u64 block_count,
                          u32* top_layer, const u64* salt_by_leaf = nullptr, u32 salt_by_leaf_cols = 0) {
  // salt_by_leaf (reference device layout, compat.cuh): column-major [salt cols][N] indexed by LEAF; gathered per coset
  // block after that block's rows are written (the rows of block 0 may alias the coefficients, which must be consumed first)
  // leaves_d row 0 is global leaf block_first * n (a shard holds only its own coset blocks' rows)
  const u64 n = (u64)1 << k, N = n << rate_bits;
  const u32 log_N = k + rate_bits;
  if (cap_height > log_N)
    return fail(P2B_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", cap_height, log_N);
  P2B_TRY(ensure_twiddles(c, log_N > 0 ? log_N - 1 : 0));
  merkle::TreeShape shape = merkle::make_shape(log_N, cap_height);
  ntt::LevelScale sc = lde_scale(k);
  const u64 leaf0 = block_first * n, nleaves = block_count * n;
  if (salt_d) {
    scatter_salt_kernel<<<(unsigned)((nleaves + 255) / 256), 256, 0, c->stream>>>(salt_d, N, log_N, leaf0, nleaves, leaves_d,
                                                                                 leaf_len, P);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  // Leaf hashing is launched per run of `hash_run` consecutive coset blocks, at least HASH_LAUNCH_MIN_LEAVES leaves: one
  // block of a small batch (2^16 rows = 512 CTAs) fills half of the 148 x 7 resident CTAs and the launches do not overlap
  // each other (measured: 2^16 x 135 15.0 -> 10.1 ms per commit, 2^17 x 234 36.5 -> 30.9 ms).
  const u64 hash_run = std::max<u64>(1, std::min<u64>(block_count, HASH_LAUNCH_MIN_LEAVES / n));
  for (u64 i = 0; i < block_count; i++) {
    u64 b = block_first + (descending ? block_count - 1 - i : i);
    if (b == 0 && wait_before_block0) CUDA_TRY(cudaStreamSynchronize(wait_before_block0));
    P2B_TRY(run_lde_block(c, c->stream, coeffs_d, n, tmp, k, P, b, sc, leaves_d, leaf_len, 0, (b - block_first) * n));
    if (salt_by_leaf) {
      gather_salt_by_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(salt_by_leaf, N, b * n, n, salt_by_leaf_cols,
                                                                                    leaves_d + (b - block_first) * n * leaf_len, leaf_len, P);
      c->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    if ((i + 1) % hash_run && i + 1 != block_count) continue;
    // hash the rows of this run of blocks on stream2 while the next block's NTT runs on stream
    const u64 run = (i % hash_run) + 1;                          // blocks in this run: the last `run` processed
    const u64 b_lo = descending ? b : b + 1 - run;               // lowest block index of the run (its rows are contiguous)
    CUDA_TRY(cudaEventRecord(c->ev_a, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_a, 0));
    P2B_TRY(launch_hash_leaves(c, c->stream2, leaves_d + (b_lo - block_first) * n * leaf_len, leaf_len, 1, (u32)leaf_len, b_lo * n,
                               run * n, shape, digests_d, cap_d));
  }
  CUDA_TRY(cudaEventRecord(c->ev_b, c->stream2));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  P2B_TRY(launch_layers(c, c->stream, shape, digests_d, cap_d, leaf0, leaf0 + nleaves, 0, top_layer));
  return P2B_OK;
}

static int lde_and_absorb_group(p2b_batch* b, u64 col0, u64 ncols, bool hash_on_stream2);

static int commit_impl(p2b_ctx* c, const u64* input, int on_host, bool is_values, u32 k, u64 P, u32 rate_bits,
                       u32 cap_height, const u64* salt, int salt_on_host, p2b_batch** out, u64 block_first = 0,
                       u64 block_count = ~(u64)0, u64* coeffs_host_out = nullptr) {
  if (!c || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (!input || P == 0) return fail(P2B_ERR_INVALID, "empty batch (no polynomials)");
  if (k + rate_bits > 32) return fail(P2B_ERR_INVALID, "degree_log + rate_bits = %u exceeds the field's two-adicity 32", k + rate_bits);
  if (P > 65535) return fail(P2B_ERR_INVALID, "more than 65535 polynomials");
  const u64 n = (u64)1 << k, N = n << rate_bits;
  const u32 log_N = k + rate_bits;
  if (cap_height > log_N)
    return fail(P2B_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", cap_height, log_N);
  CUDA_TRY(cudaSetDevice(c->device));
  const u64 salt_size = salt ? P2B_SALT_SIZE : 0, leaf_len = P + salt_size;
  const u64 ncap = (u64)1 << cap_height, ndig = 2 * (N - ncap);
  const u64 R = (u64)1 << rate_bits;
  if (block_count == ~(u64)0) block_count = R - block_first;
  if (block_first >= R || block_count == 0 || block_first + block_count > R)
    return fail(P2B_ERR_INVALID, "coset block range [%llu, +%llu) outside 2^rate_bits = %llu", (unsigned long long)block_first,
                (unsigned long long)block_count, (unsigned long long)R);

  p2b_batch* b = new (std::nothrow) p2b_batch();
  if (!b) return fail(P2B_ERR_OOM, "host allocation failed");
  b->ctx = c;
  b->info = p2b_batch_info{k, rate_bits, cap_height, (u32)salt_size, P, N, leaf_len, ndig};
  b->shape = merkle::make_shape(log_N, cap_height);
  b->first_leaf = block_first * n;
  b->local_leaves = block_count * n;
  cudaStream_t st = c->stream;
  u64* staged = nullptr;  // device copy of host input
  u64* salt_d = nullptr;
  int rc = P2B_OK;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&b->coeffs, P * n * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->leaves, b->local_leaves * leaf_len * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->digests, (ndig ? ndig : 1) * 4 * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->cap, ncap * 4 * sizeof(u64), st));
    P2B_TRY(ensure_scratch(c, P * n));
    const u64* in_d = input;
    bool ifft_done = false;
    if (on_host && is_values && !salt && P >= 16 && block_first == 0 && block_count == R) {
      // Host values, whole batch: everything is pipelined behind the upload.  Groups of 16 columns (two sponge blocks):
      // H2D on the copy stream -> inverse NTT -> (coefficients back to the host on the other DMA engine) -> LDE into the
      // leaf rows -> the leaves' sponges absorb the group on stream2 while the next group uploads and transforms.
      // Only the first group's upload is exposed (the one-shot flow waits for the whole matrix before it can hash).
      // groups of 8, 8, 16, 16, ... columns: the first upload, the only exposed one, is half as long
      auto group_width = [](u64 g) -> u64 { return g < 2 ? 8 : 16; };
      CUDA_TRY(pool_alloc(&b->sponge_state, 12 * b->local_leaves * sizeof(u64), st));
      P2B_TRY(ensure_twiddles(c, log_N > 0 ? log_N - 1 : 0));
      CUDA_TRY(cudaEventRecord(c->ev_copy[31], st));                 // allocations above are stream-ordered on st
      CUDA_TRY(cudaStreamWaitEvent(c->stream_h2d, c->ev_copy[31], 0));
      for (u64 g = 0, c0 = 0; c0 < P; c0 += group_width(g), g++) {
        const u64 c1 = std::min<u64>(P, c0 + group_width(g));
        cudaEvent_t ev_up = c->ev_copy[g % 14], ev_dn = c->ev_copy[14 + g % 14];
        CUDA_TRY(cudaMemcpyAsync(b->coeffs + c0 * n, input + c0 * n, (c1 - c0) * n * sizeof(u64), cudaMemcpyHostToDevice, c->stream_h2d));
        CUDA_TRY(cudaEventRecord(ev_up, c->stream_h2d));
        CUDA_TRY(cudaStreamWaitEvent(st, ev_up, 0));
        P2B_TRY(run_ifft(c, b->coeffs + c0 * n, b->coeffs + c0 * n, c->scratch + c0 * n, k, c1 - c0));
        if (coeffs_host_out) {
          CUDA_TRY(cudaEventRecord(ev_dn, st));
          CUDA_TRY(cudaStreamWaitEvent(c->stream_d2h, ev_dn, 0));
          CUDA_TRY(cudaMemcpyAsync(coeffs_host_out + c0 * n, b->coeffs + c0 * n, (c1 - c0) * n * sizeof(u64), cudaMemcpyDeviceToHost, c->stream_d2h));
        }
        P2B_TRY(lde_and_absorb_group(b, c0, c1 - c0, true));
      }
      CUDA_TRY(cudaEventRecord(c->ev_b, c->stream2));
      CUDA_TRY(cudaStreamWaitEvent(st, c->ev_b, 0));
      P2B_TRY(launch_layers(c, st, b->shape, b->digests, b->cap, b->first_leaf, b->first_leaf + b->local_leaves, 0, &b->top_layer));
      if (coeffs_host_out) {
        CUDA_TRY(cudaEventRecord(c->ev_copy[29], c->stream_d2h));
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy[29], 0));      // a later synchronisation of the main stream covers the copies
      }
      CUDA_TRY(cudaFreeAsync(b->sponge_state, st));
      b->sponge_state = nullptr;
      return P2B_OK;
    }
    if (on_host && is_values && P >= 16) {
      // Host values: H2D in column groups on the copy stream, each group's inverse NTT starts as soon as its columns
      // have landed (the transform is per column), so only the last group's transform is exposed after the copy.
      const u32 groups = 16;
      CUDA_TRY(cudaEventRecord(c->ev_copy[31], st));                 // allocations above are stream-ordered on st
      CUDA_TRY(cudaStreamWaitEvent(c->stream_h2d, c->ev_copy[31], 0));
      for (u32 g = 0; g < groups; g++) {
        u64 c0 = P * g / groups, c1 = P * (g + 1) / groups;
        if (c1 == c0) continue;
        CUDA_TRY(cudaMemcpyAsync(b->coeffs + c0 * n, input + c0 * n, (c1 - c0) * n * sizeof(u64), cudaMemcpyHostToDevice, c->stream_h2d));
        CUDA_TRY(cudaEventRecord(c->ev_copy[g], c->stream_h2d));
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy[g], 0));
        P2B_TRY(run_ifft(c, b->coeffs + c0 * n, b->coeffs + c0 * n, c->scratch + c0 * n, k, c1 - c0));
      }
      in_d = b->coeffs;
      ifft_done = true;
    } else if (on_host) {
      // host input goes straight into the coefficient buffer (values are transformed in place there)
      CUDA_TRY(cudaMemcpyAsync(b->coeffs, input, P * n * sizeof(u64), cudaMemcpyHostToDevice, st));
      in_d = b->coeffs;
    }
    if (salt) {
      if (salt_on_host) {
        CUDA_TRY(pool_alloc(&salt_d, P2B_SALT_SIZE * N * sizeof(u64), st));
        CUDA_TRY(cudaMemcpyAsync(salt_d, salt, P2B_SALT_SIZE * N * sizeof(u64), cudaMemcpyHostToDevice, st));
      } else {
        salt_d = const_cast<u64*>(salt);
      }
    }
    if (is_values) {
      if (!ifft_done) P2B_TRY(run_ifft(c, in_d, b->coeffs, c->scratch, k, P));
    } else if (!on_host) {
      CUDA_TRY(cudaMemcpyAsync(b->coeffs, input, P * n * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    }
    if (coeffs_host_out) {
      // coefficients back to the host (the reference keeps `polynomials` host-side, oracle.rs:403-407) on the D2H
      // engine while the LDE and the Merkle tree are computed
      CUDA_TRY(cudaEventRecord(c->ev_copy[30], st));
      CUDA_TRY(cudaStreamWaitEvent(c->stream_d2h, c->ev_copy[30], 0));
      CUDA_TRY(cudaMemcpyAsync(coeffs_host_out, b->coeffs, P * n * sizeof(u64), cudaMemcpyDeviceToHost, c->stream_d2h));
      CUDA_TRY(cudaEventRecord(c->ev_copy[29], c->stream_d2h));
    }
    P2B_TRY(lde_and_merkle(c, b->coeffs, k, P, rate_bits, cap_height, salt_d, c->scratch, b->leaves, leaf_len,
                           b->digests, b->cap, false, nullptr, block_first, block_count, &b->top_layer));
    // a later synchronisation of the main stream also covers the coefficient copy
    if (coeffs_host_out) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy[29], 0));
    return P2B_OK;
  };
  rc = body();
  if (salt_d && salt_on_host) cudaFreeAsync(salt_d, st);
  if (staged) cudaFreeAsync(staged, st);
  if (rc != P2B_OK) {
    batch_free(b);
    return rc;
  }
  *out = b;
  return P2B_OK;
}

extern "C" int p2b_commit_from_values(p2b_ctx* ctx, const uint64_t* values, int values_on_host, uint32_t n_log,
                                      uint64_t P, uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt,
                                      int salt_on_host, p2b_batch** out) {
  return commit_impl(ctx, values, values_on_host, true, n_log, P, rate_bits, cap_height, salt, salt_on_host, out);
}
extern "C" int p2b_commit_from_values_ex(p2b_ctx* ctx, const uint64_t* values, int values_on_host, uint32_t n_log, uint64_t P,
                                         uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, int salt_on_host,
                                         uint64_t* coeffs_host_out, p2b_batch** out) {
  return commit_impl(ctx, values, values_on_host, true, n_log, P, rate_bits, cap_height, salt, salt_on_host, out, 0, ~(u64)0,
                     coeffs_host_out);
}
extern "C" int p2b_commit_from_coeffs(p2b_ctx* ctx, const uint64_t* coeffs, int coeffs_on_host, uint32_t n_log,
                                      uint64_t P, uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt,
                                      int salt_on_host, p2b_batch** out) {
  return commit_impl(ctx, coeffs, coeffs_on_host, false, n_log, P, rate_bits, cap_height, salt, salt_on_host, out);
}

// ---- sharded commit (multi-GPU): one rank = a contiguous range of coset blocks ------------------------------
extern "C" int p2b_commit_blocks(p2b_ctx* ctx, const uint64_t* d_coeffs, uint32_t n_log, uint64_t P, uint32_t rate_bits,
                                 uint32_t cap_height, const uint64_t* d_salt, uint64_t block_first, uint64_t block_count,
                                 p2b_batch** out) {
  return commit_impl(ctx, d_coeffs, 0, false, n_log, P, rate_bits, cap_height, d_salt, 0, out, block_first, block_count);
}
// ======================================================================================================
// pipelined commit of coset blocks: coefficient columns arrive in groups (one all-gather round each in the multi-GPU
// flow); each group is low-degree-extended into the leaf rows and absorbed by the leaves' sponges at once, so the
// exchange of group g+1 overlaps the LDE + hashing of group g.  Same results as p2b_commit_blocks.
// ======================================================================================================
extern "C" int p2b_commit_blocks_begin(p2b_ctx* c, uint32_t k, uint64_t P, uint32_t rate_bits, uint32_t cap_height,
                                       uint64_t block_first, uint64_t block_count, p2b_batch** out) {
  if (!c || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (P == 0) return fail(P2B_ERR_INVALID, "empty batch (no polynomials)");
  if (P <= 4) return fail(P2B_ERR_UNSUPPORTED, "pipelined commit needs more than 4 polynomials (hash_or_noop copies shorter leaves)");
  if (k + rate_bits > 32) return fail(P2B_ERR_INVALID, "degree_log + rate_bits = %u exceeds the field's two-adicity 32", k + rate_bits);
  if (P > 65535) return fail(P2B_ERR_INVALID, "more than 65535 polynomials");
  const u64 n = (u64)1 << k, N = n << rate_bits, R = (u64)1 << rate_bits;
  const u32 log_N = k + rate_bits;
  if (cap_height > log_N) return fail(P2B_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", cap_height, log_N);
  if (block_first >= R || block_count == 0 || block_first + block_count > R)
    return fail(P2B_ERR_INVALID, "coset block range [%llu, +%llu) outside 2^rate_bits = %llu", (unsigned long long)block_first,
                (unsigned long long)block_count, (unsigned long long)R);
  CUDA_TRY(cudaSetDevice(c->device));
  const u64 ncap = (u64)1 << cap_height, ndig = 2 * (N - ncap);
  p2b_batch* b = new (std::nothrow) p2b_batch();
  if (!b) return fail(P2B_ERR_OOM, "host allocation failed");
  b->ctx = c;
  b->info = p2b_batch_info{k, rate_bits, cap_height, 0, P, N, P, ndig};
  b->shape = merkle::make_shape(log_N, cap_height);
  b->first_leaf = block_first * n;
  b->local_leaves = block_count * n;
  b->pipelined = true;
  cudaStream_t st = c->stream;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&b->coeffs, P * n * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->leaves, b->local_leaves * P * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->digests, (ndig ? ndig : 1) * 4 * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->cap, ncap * 4 * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&b->sponge_state, 12 * b->local_leaves * sizeof(u64), st));
    P2B_TRY(ensure_twiddles(c, log_N > 0 ? log_N - 1 : 0));
    return P2B_OK;
  };
  int rc = body();
  if (rc != P2B_OK) {
    batch_free(b);
    return rc;
  }
  *out = b;
  return P2B_OK;
}

// LDE of coefficient columns [col0, col0 + ncols) (already in b->coeffs) into the leaf rows of the batch's coset blocks on
// the main stream, then the leaves' sponges absorb them -- on stream2 when `hash_on_stream2` (the caller's next group then
// transforms on the main stream meanwhile; absorbs stay ordered among themselves on stream2).
static int lde_and_absorb_group(p2b_batch* b, u64 col0, u64 ncols, bool hash_on_stream2) {
  p2b_ctx* c = b->ctx;
  cudaStream_t st = c->stream;
  const u64 P = b->info.num_polys;
  const u32 k = b->info.degree_log;
  const u64 n = (u64)1 << k;
  const bool last = col0 + ncols == P;
  P2B_TRY(ensure_scratch(c, ncols * n));
  ntt::LevelScale sc = lde_scale(k);
  const u64 block_first = b->first_leaf / n, block_count = b->local_leaves / n;
  for (u64 i = 0; i < block_count; i++)
    P2B_TRY(run_lde_block(c, st, b->coeffs + col0 * n, n, c->scratch, k, ncols, block_first + i, sc, b->leaves, P, col0, i * n));
  cudaStream_t hs = st;
  if (hash_on_stream2) {
    CUDA_TRY(cudaEventRecord(c->ev_a, st));
    CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_a, 0));
    hs = c->stream2;
  }
  const u64 count = b->local_leaves;
  Stage stage(c, hs, "build Merkle tree: leaf hashes");
  unsigned blocks = (unsigned)((count + P2B_HASH_BLOCK - 1) / P2B_HASH_BLOCK);
  std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
  if (c->time_hash) {
    if (c->hash_events_used == c->hash_events.size()) {
      cudaEvent_t a, e2;
      CUDA_TRY(cudaEventCreate(&a));
      CUDA_TRY(cudaEventCreate(&e2));
      c->hash_events.emplace_back(a, e2);
    }
    ev = &c->hash_events[c->hash_events_used++];
    CUDA_TRY(cudaEventRecord(ev->first, hs));
  }
  merkle::absorb_columns_kernel<<<blocks, P2B_HASH_BLOCK, 0, hs>>>(b->leaves, P, (u32)col0, (u32)ncols, last ? 1 : 0, count, b->first_leaf,
                                                                   b->shape, b->sponge_state, b->digests, b->cap);
  if (ev) CUDA_TRY(cudaEventRecord(ev->second, hs));
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  b->cols_done = col0 + ncols;
  return P2B_OK;
}

extern "C" int p2b_commit_blocks_absorb(p2b_batch* b, const uint64_t* d_coeff_cols, uint64_t col0, uint64_t ncols) {
  if (!b || !d_coeff_cols) return fail(P2B_ERR_INVALID, "NULL argument");
  if (!b->pipelined) return fail(P2B_ERR_INVALID, "batch was not created by p2b_commit_blocks_begin");
  const u64 P = b->info.num_polys;
  if (col0 != b->cols_done) return fail(P2B_ERR_INVALID, "columns must be absorbed in order: expected column %llu, got %llu",
                                        (unsigned long long)b->cols_done, (unsigned long long)col0);
  if (ncols == 0 || col0 + ncols > P) return fail(P2B_ERR_INVALID, "column range [%llu, +%llu) outside the %llu polynomials",
                                                  (unsigned long long)col0, (unsigned long long)ncols, (unsigned long long)P);
  const bool last = col0 + ncols == P;
  if (!last && (ncols % 8)) return fail(P2B_ERR_INVALID, "every group but the last must hold a multiple of 8 columns (sponge rate)");
  p2b_ctx* c = b->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const u64 n = (u64)1 << b->info.degree_log;
  u64* dst = b->coeffs + col0 * n;
  if (dst != d_coeff_cols) CUDA_TRY(cudaMemcpyAsync(dst, d_coeff_cols, ncols * n * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
  return lde_and_absorb_group(b, col0, ncols, false);
}

extern "C" int p2b_commit_blocks_finish(p2b_batch* b) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  if (!b->pipelined) return fail(P2B_ERR_INVALID, "batch was not created by p2b_commit_blocks_begin");
  if (b->cols_done != b->info.num_polys)
    return fail(P2B_ERR_INVALID, "only %llu of %llu columns absorbed", (unsigned long long)b->cols_done, (unsigned long long)b->info.num_polys);
  p2b_ctx* c = b->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  P2B_TRY(launch_layers(c, c->stream, b->shape, b->digests, b->cap, b->first_leaf, b->first_leaf + b->local_leaves, 0, &b->top_layer));
  if (b->sponge_state) {
    cudaFreeAsync(b->sponge_state, c->stream);
    b->sponge_state = nullptr;
  }
  b->pipelined = false;
  return P2B_OK;
}

extern "C" int p2b_batch_shard_info(const p2b_batch* b, uint64_t* first_leaf, uint64_t* local_leaves, uint32_t* top_layer,
                                    uint64_t* top_node_first, uint64_t* top_node_count) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  if (first_leaf) *first_leaf = b->first_leaf;
  if (local_leaves) *local_leaves = b->local_leaves;
  if (top_layer) *top_layer = b->top_layer;
  if (top_node_first) *top_node_first = b->first_leaf >> b->top_layer;
  if (top_node_count) *top_node_count = b->local_leaves >> b->top_layer;
  return P2B_OK;
}
// nodes [node_first, +count) of digest layer `layer` <-> contiguous device buffer [count][4]
__global__ void move_nodes_kernel(merkle::TreeShape shape, u32 layer, u64 node_first, u64 count, u64* digests, u64* cap,
                                  u64* buf, int import) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count * 4) return;
  u64* slot = merkle::node_slot(shape, digests, cap, layer, node_first + (i >> 2)) + (i & 3);
  if (import) *slot = buf[i];
  else buf[i] = *slot;
}
static int move_nodes(const p2b_batch* b, u32 layer, u64 node_first, u64 count, u64* d_buf, int import) {
  if (!b || !d_buf) return fail(P2B_ERR_INVALID, "NULL argument");
  if (layer > b->shape.sub_log || node_first + count > (b->info.num_leaves >> layer)) return fail(P2B_ERR_INVALID, "node range out of bounds");
  if (count == 0) return P2B_OK;
  p2b_ctx* c = b->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  move_nodes_kernel<<<(unsigned)((count * 4 + 255) / 256), 256, 0, c->stream>>>(b->shape, layer, node_first, count, b->digests, b->cap,
                                                                          d_buf, import);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}
extern "C" int p2b_batch_export_nodes(const p2b_batch* b, uint32_t layer, uint64_t node_first, uint64_t count, uint64_t* d_out) {
  return move_nodes(b, layer, node_first, count, d_out, 0);
}
extern "C" int p2b_batch_import_nodes(p2b_batch* b, uint32_t layer, uint64_t node_first, uint64_t count, const uint64_t* d_in) {
  return move_nodes(b, layer, node_first, count, const_cast<u64*>(d_in), 1);
}
// after every shard's top-layer nodes have been imported: the remaining layers up to the cap (whole tree)
extern "C" int p2b_batch_finish_layers(p2b_batch* b, uint32_t from_layer) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  if (from_layer > b->shape.sub_log) return fail(P2B_ERR_INVALID, "layer out of range");
  p2b_ctx* c = b->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  u32 top = 0;
  P2B_TRY(launch_layers(c, c->stream, b->shape, b->digests, b->cap, 0, b->info.num_leaves, from_layer, &top));
  b->top_layer = top;
  return P2B_OK;
}

extern "C" int p2b_batch_get_info(const p2b_batch* b, p2b_batch_info* out) {
  if (!b || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = b->info;
  return P2B_OK;
}
extern "C" int p2b_batch_device_ptrs(const p2b_batch* b, uint64_t** coeffs, uint64_t** leaves, uint64_t** digests,
                                     uint64_t** cap) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  if (coeffs) *coeffs = b->coeffs;
  if (leaves) *leaves = b->leaves;
  if (digests) *digests = b->digests;
  if (cap) *cap = b->cap;
  return P2B_OK;
}

static int d2h(const p2b_batch* b, void* dst, const void* src, size_t bytes) {
  if (!dst) return fail(P2B_ERR_INVALID, "NULL output");
  if (bytes == 0) return P2B_OK;
  CUDA_TRY(cudaSetDevice(b->ctx->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(b->ctx->stream));
  return P2B_OK;
}
extern "C" int p2b_batch_get_coeffs(const p2b_batch* b, uint64_t* out) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  return d2h(b, out, b->coeffs, (b->info.num_polys << b->info.degree_log) * sizeof(u64));
}
extern "C" int p2b_batch_get_cap(const p2b_batch* b, uint64_t* out) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  return d2h(b, out, b->cap, ((size_t)4 << b->info.cap_height) * sizeof(u64));
}
extern "C" int p2b_batch_get_digests(const p2b_batch* b, uint64_t* out) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  return d2h(b, out, b->digests, b->info.num_digests * 4 * sizeof(u64));
}
extern "C" int p2b_batch_get_leaves(const p2b_batch* b, uint64_t first_leaf, uint64_t count, uint64_t* out) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  if (first_leaf < b->first_leaf || first_leaf + count > b->first_leaf + b->local_leaves)
    return fail(P2B_ERR_INVALID, "leaf range out of bounds (this batch holds leaves [%llu, %llu))", (unsigned long long)b->first_leaf,
                (unsigned long long)(b->first_leaf + b->local_leaves));
  return d2h(b, out, b->leaves + (first_leaf - b->first_leaf) * b->info.leaf_len, count * b->info.leaf_len * sizeof(u64));
}
extern "C" int p2b_batch_get_lde_values(const p2b_batch* b, uint64_t index, uint64_t step, uint64_t* out) {
  if (!b) return fail(P2B_ERR_INVALID, "NULL batch");
  u32 bits = b->info.degree_log + b->info.rate_bits;
  u64 i = index * step;
  if (i >= b->info.num_leaves) return fail(P2B_ERR_INVALID, "index*step out of the LDE domain");
  u64 row = 0;
  for (u32 t = 0; t < bits; t++) row |= ((i >> t) & 1) << (bits - 1 - t);  // reverse_bits, util/mod.rs:55-63
  if (row < b->first_leaf || row >= b->first_leaf + b->local_leaves)
    return fail(P2B_ERR_INVALID, "row %llu is held by another shard", (unsigned long long)row);
  return d2h(b, out, b->leaves + (row - b->first_leaf) * b->info.leaf_len, b->info.num_polys * sizeof(u64));
}

// gather rows and/or Merkle paths for a list of leaf indices
__global__ void open_rows_kernel(const u64* __restrict__ leaves, u64 leaf_len, const u64* __restrict__ digests,
                                 merkle::TreeShape shape, const u64* __restrict__ idx, u64 count, u64* __restrict__ rows,
                                 u64* __restrict__ sibs) {
  u64 qi = blockIdx.x;
  if (qi >= count) return;
  u64 leaf = idx[qi];
  if (rows)
    for (u64 j = threadIdx.x; j < leaf_len; j += blockDim.x) rows[qi * leaf_len + j] = leaves[leaf * leaf_len + j];
  if (sibs) {
    // MerkleTree::prove, merkle_tree.rs:392-440
    u32 layers = shape.sub_log;
    u64 tree = leaf >> layers;
    const u64* dt = digests + 4 * tree * shape.sub_digests;
    for (u32 t = threadIdx.x; t < layers * 4; t += blockDim.x) {
      u32 i = t >> 2, w = t & 3;
      u64 pair_index = (leaf & (((u64)1 << layers) - 1)) >> i;
      u64 parity = pair_index & 1;
      pair_index >>= 1;
      u64 siblings_index = (pair_index << (i + 1)) + ((u64)1 << i) - 1;
      u64 sibling_index = 2 * siblings_index + (1 - parity);
      sibs[(qi * layers + i) * 4 + w] = dt[4 * sibling_index + w];
    }
  }
}

static int open_impl(const p2b_batch* b, const u64* leaf_indices, u64 count, u64* rows_out, u64* sibs_out) {
  if (!b || !leaf_indices) return fail(P2B_ERR_INVALID, "NULL argument");
  if (count == 0) return P2B_OK;
  for (u64 i = 0; i < count; i++)
    if (leaf_indices[i] < b->first_leaf || leaf_indices[i] >= b->first_leaf + b->local_leaves)
      return fail(P2B_ERR_INVALID, "leaf index %llu out of range (this batch holds leaves [%llu, %llu))", (unsigned long long)leaf_indices[i],
                  (unsigned long long)b->first_leaf, (unsigned long long)(b->first_leaf + b->local_leaves));
  p2b_ctx* c = b->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const u64 ll = b->info.leaf_len, layers = b->shape.sub_log;
  u64 *d_idx = nullptr, *d_rows = nullptr, *d_sibs = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&d_idx, count * sizeof(u64), st));
    CUDA_TRY(cudaMemcpyAsync(d_idx, leaf_indices, count * sizeof(u64), cudaMemcpyHostToDevice, st));
    if (rows_out) CUDA_TRY(pool_alloc(&d_rows, count * ll * sizeof(u64), st));
    if (sibs_out && layers) CUDA_TRY(pool_alloc(&d_sibs, count * layers * 4 * sizeof(u64), st));
    open_rows_kernel<<<(unsigned)count, 128, 0, st>>>(b->leaves - b->first_leaf * ll, ll, b->digests, b->shape, d_idx, count, d_rows, d_sibs);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if (d_rows) CUDA_TRY(cudaMemcpyAsync(rows_out, d_rows, count * ll * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (d_sibs) CUDA_TRY(cudaMemcpyAsync(sibs_out, d_sibs, count * layers * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return P2B_OK;
  };
  int rc = body();
  if (d_idx) cudaFreeAsync(d_idx, st);
  if (d_rows) cudaFreeAsync(d_rows, st);
  if (d_sibs) cudaFreeAsync(d_sibs, st);
  return rc;
}
extern "C" int p2b_batch_prove(const p2b_batch* b, const uint64_t* leaf_indices, uint64_t count, uint64_t* siblings_out) {
  if (!siblings_out && b && b->shape.sub_log) return fail(P2B_ERR_INVALID, "NULL output");
  return open_impl(b, leaf_indices, count, nullptr, siblings_out);
}
extern "C" int p2b_batch_open_rows(const p2b_batch* b, const uint64_t* leaf_indices, uint64_t count, uint64_t* rows_out,
                                   uint64_t* siblings_out) {
  if (!rows_out) return fail(P2B_ERR_INVALID, "NULL output");
  return open_impl(b, leaf_indices, count, rows_out, siblings_out);
}

// ======================================================================================================
// building blocks
// ======================================================================================================
extern "C" int p2b_ifft_batch(p2b_ctx* c, const uint64_t* d_src, uint64_t* d_dst, uint32_t n_log, uint64_t P) {
  if (!c || !d_src || !d_dst) return fail(P2B_ERR_INVALID, "NULL argument");
  if (n_log > 32) return fail(P2B_ERR_INVALID, "n_log exceeds two-adicity");
  if (P > 65535) return fail(P2B_ERR_INVALID, "more than 65535 polynomials");
  CUDA_TRY(cudaSetDevice(c->device));
  Plan pl = make_plan(n_log);
  u64* tmp = d_dst;
  if (pl.n_strided) {
    P2B_TRY(ensure_scratch(c, P << n_log));
    tmp = c->scratch;
  }
  return run_ifft(c, d_src, d_dst, tmp, n_log, P);
}

extern "C" int p2b_lde_leaves(p2b_ctx* c, const uint64_t* d_coeffs, uint32_t n_log, uint64_t P, uint32_t rate_bits,
                              uint64_t* d_leaves, uint64_t row_stride, uint64_t col0) {
  if (!c || !d_coeffs || !d_leaves) return fail(P2B_ERR_INVALID, "NULL argument");
  if (n_log + rate_bits > 32) return fail(P2B_ERR_INVALID, "degree_log + rate_bits exceeds two-adicity");
  if (P > 65535) return fail(P2B_ERR_INVALID, "more than 65535 polynomials");
  if (col0 + P > row_stride) return fail(P2B_ERR_INVALID, "columns do not fit the row stride");
  CUDA_TRY(cudaSetDevice(c->device));
  u32 log_N = n_log + rate_bits;
  P2B_TRY(ensure_twiddles(c, log_N > 0 ? log_N - 1 : 0));
  P2B_TRY(ensure_scratch(c, P << n_log));
  ntt::LevelScale sc = lde_scale(n_log);
  for (u64 b = 0; b < ((u64)1 << rate_bits); b++)
    P2B_TRY(run_lde_block(c, c->stream, d_coeffs, (u64)1 << n_log, c->scratch, n_log, P, b, sc, d_leaves, row_stride, col0,
                          b << n_log));
  return P2B_OK;
}

extern "C" int p2b_merkle_tree(p2b_ctx* c, const uint64_t* d_leaves, uint64_t num_leaves, uint64_t leaf_len,
                               uint64_t row_stride, uint64_t col_stride, uint32_t cap_height, uint64_t* d_digests,
                               uint64_t* d_cap) {
  if (!c || !d_leaves || !d_cap) return fail(P2B_ERR_INVALID, "NULL argument");
  if (num_leaves == 0 || (num_leaves & (num_leaves - 1))) return fail(P2B_ERR_INVALID, "number of leaves must be a power of two");
  u32 lg = 0;
  while (((u64)1 << lg) < num_leaves) lg++;
  if (cap_height > lg) return fail(P2B_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", cap_height, lg);
  if (cap_height < lg && !d_digests) return fail(P2B_ERR_INVALID, "NULL digests");
  if (leaf_len > 0xffffffffull) return fail(P2B_ERR_INVALID, "leaf too long");
  CUDA_TRY(cudaSetDevice(c->device));
  merkle::TreeShape shape = merkle::make_shape(lg, cap_height);
  P2B_TRY(launch_hash_leaves(c, c->stream, d_leaves, row_stride, col_stride, (u32)leaf_len, 0, num_leaves, shape, d_digests, d_cap));
  return launch_layers(c, c->stream, shape, d_digests, d_cap, 0, num_leaves, 0, nullptr);
}

extern "C" int p2b_poseidon_permute(p2b_ctx* c, uint64_t* d_states, uint64_t count) {
  if (!c || !d_states) return fail(P2B_ERR_INVALID, "NULL argument");
  if (count == 0) return P2B_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  merkle::permute_kernel<<<(unsigned)((count + P2B_HASH_BLOCK - 1) / P2B_HASH_BLOCK), P2B_HASH_BLOCK, 0, c->stream>>>(d_states, count, 1);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

__global__ void field_op_kernel(int op, const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out, u64 count) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 x = a[i], y = b[i], r;
  switch (op) {
    case 0: r = gl::add(x, y); break;
    case 1: r = gl::sub(x, y); break;
    case 2: r = gl::mul(x, y); break;
    case 3: r = gl::mul_add(x, y, x); break;       // x*y + x
    case 4: r = gl::add_canonical(x, gl::canon(y)); break;
    case 5: r = gl::sub_canonical(x, gl::canon(y)); break;
    case 6: case 7: {   // the optimistic add / sub of the NTT butterflies, with the caller's duty: exact redo when flagged
      gl::Optimistic m;
      r = op == 6 ? gl::add(x, y, m) : gl::sub(x, y, m);
      if (m.any()) r = op == 6 ? gl::add(x, y) : gl::sub(x, y);
      break;
    }
    case 8: case 9: {   // the flag alone: 1 where the single repayment wrapped again
      gl::Optimistic m;
      (void)(op == 8 ? gl::add(x, y, m) : gl::sub(x, y, m));
      out[i] = m.any() ? 1 : 0;
      return;
    }
    default: r = 0;
  }
  out[i] = gl::canon(r);
}
extern "C" int p2b_field_op(p2b_ctx* c, int op, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint64_t count) {
  if (!c || !d_a || !d_b || !d_out) return fail(P2B_ERR_INVALID, "NULL argument");
  if (op < 0 || op > 9) return fail(P2B_ERR_INVALID, "unknown op %d", op);
  if (count == 0) return P2B_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  field_op_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(op, d_a, d_b, d_out, count);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

// splitmix64(seed + index) with rejection of values >= p (re-mix until canonical), BASELINE.md C2
__host__ __device__ static inline u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void fill_synthetic_kernel(u64* __restrict__ out, u64 count, u64 seed, u64 first) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 v = splitmix64(seed ^ splitmix64(first + i));
  while (v >= gl::P) v = splitmix64(v);
  out[i] = v;
}
extern "C" int p2b_fill_synthetic(p2b_ctx* c, uint64_t* d_out, uint64_t count, uint64_t seed, uint64_t first_index) {
  if (!c || !d_out) return fail(P2B_ERR_INVALID, "NULL argument");
  if (count == 0) return P2B_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  fill_synthetic_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(d_out, count, seed, first_index);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

extern "C" int p2b_malloc(p2b_ctx* c, uint64_t bytes, void** out) {
  if (!c || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  // From the device's stream-ordered pool (release threshold = never, ctx_init_common), ordered on the context's stream:
  // after the first proof a buffer of a prove() stage costs microseconds instead of a cudaMalloc / cudaFree pair (~1 ms per
  // 100 MB, and cudaFree synchronises the device).  Every library call starts on that stream or makes its other streams
  // wait for it first, and joins them back before it returns, so stream order on it covers all uses of the buffer.
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(pool_alloc(out, bytes ? bytes : 1, c->stream));
  return P2B_OK;
}
extern "C" int p2b_free(p2b_ctx* c, void* ptr) {
  if (!c) return fail(P2B_ERR_INVALID, "NULL argument");
  if (!ctx_alive(c) || !ptr) return P2B_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaFreeAsync(ptr, c->stream));
  return P2B_OK;
}
extern "C" int p2b_malloc_host(uint64_t bytes, void** out) {
  if (!out) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
  return P2B_OK;
}
extern "C" int p2b_free_host(void* ptr) {
  CUDA_TRY(cudaFreeHost(ptr));
  return P2B_OK;
}
extern "C" int p2b_memcpy_h2d(p2b_ctx* c, void* d_dst, const void* h_src, uint64_t bytes) {
  if (!c) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return P2B_OK;
}
extern "C" int p2b_memcpy_d2h(p2b_ctx* c, void* h_dst, const void* d_src, uint64_t bytes) {
  if (!c) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return P2B_OK;
}
extern "C" int p2b_timer_start(p2b_ctx* c) {
  if (!c) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaEventRecord(c->ev_t0, c->stream));
  return P2B_OK;
}
extern "C" int p2b_timer_stop_ms(p2b_ctx* c, float* ms_out) {
  if (!c || !ms_out) return fail(P2B_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaEventRecord(c->ev_t1, c->stream));
  CUDA_TRY(cudaEventSynchronize(c->ev_t1));
  CUDA_TRY(cudaEventElapsedTime(ms_out, c->ev_t0, c->ev_t1));
  return P2B_OK;
}

// ======================================================================================================
// quotient polynomials (plonky2/src/plonk/prover.rs:790-1034)
// ======================================================================================================
// coset_ifft(F::coset_shift()) per challenge (prover.rs:1014-1021, polynomial/mod.rs:64-77): values [nc][2^lde_log] -> coefficients
static int quotient_values_to_coeffs(p2b_ctx* c, const u64* vals, u64* d_coeffs_out, u32 lde_log, u32 nc) {
  cudaStream_t st = c->stream;
  const u64 lde_size = (u64)1 << lde_log;
  Plan pl = make_plan(lde_log);
  u64* tmp = d_coeffs_out;
  if (pl.n_strided) {
    P2B_TRY(ensure_scratch(c, nc * lde_size));
    tmp = c->scratch;
  }
  P2B_TRY(run_ifft(c, vals, d_coeffs_out, tmp, lde_log, nc));
  quotient::scale_by_powers_kernel<<<(unsigned)((lde_size + 255) / 256), 256, 0, st>>>(d_coeffs_out, lde_size, nc, hostf::inv(hostf::COSET_SHIFT));
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return P2B_OK;
}

static int quotient_impl(p2b_ctx* c, const p2b_circuit* circ, const u64* d_wires, u64 wires_stride, const u64* d_zs_pp,
                         u64 zs_stride, const u64* d_cs, u64 cs_stride, const u64* pih, const u64* betas, const u64* gammas,
                         const u64* alphas, u64* d_values_out, u64* d_coeffs_out, u64* d_rows_out, u64 pt_first = 0, u64 pt_stride = 1,
                         u64 last_row = ~(u64)0) {
  // pt_first / pt_stride: evaluate only the points i = pt_first + pt_stride * j (a device of a multi-device prover owns the
  // points whose rows it holds, mgpu.cuh); the row pointers are then "virtual" bases (shard pointer - first_leaf * stride)
  // and last_row the last leaf row the shard holds.  Values come out compact: [nc][lde_size / pt_stride].
  if (!c || !circ || !d_wires || !d_zs_pp || !d_cs || !pih || !betas || !gammas || !alphas) return fail(P2B_ERR_INVALID, "NULL argument");
  if (pt_stride != 1 && d_coeffs_out) return fail(P2B_ERR_INVALID, "coefficients need the values of the whole domain");
  Stage stage(c, c->stream, "compute quotient polys");   // plonk/prover.rs:187-196
  if (!circ->gates && circ->num_gates) return fail(P2B_ERR_INVALID, "NULL gate list");
  if (!circ->k_is && circ->num_routed_wires) return fail(P2B_ERR_INVALID, "NULL k_is");
  const u32 nc = circ->num_challenges, qdf = circ->quotient_degree_factor;
  if (nc == 0 || nc > quotient::MAX_CHALLENGES) return fail(P2B_ERR_INVALID, "num_challenges must be 1..%d", quotient::MAX_CHALLENGES);
  if (qdf < 2) return fail(P2B_ERR_INVALID, "quotient_degree_factor must be at least 2");
  u32 qdb = 0;
  while ((1u << qdb) < qdf) qdb++;  // log2_ceil
  if (qdb > circ->rate_bits)
    return fail(P2B_ERR_UNSUPPORTED, "constraints of degree higher than the rate are not supported (prover.rs:809-813)");
  if ((1u << qdb) > quotient::MAX_ZH) return fail(P2B_ERR_UNSUPPORTED, "quotient degree factor too large");
  if (circ->degree_bits + circ->rate_bits > 32) return fail(P2B_ERR_INVALID, "degree_bits + rate_bits exceeds two-adicity");
  if (circ->num_selectors > circ->num_constants) return fail(P2B_ERR_INVALID, "more selectors than constants");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;

  quotient::Params p{};
  p.degree_bits = circ->degree_bits;
  p.rate_bits = circ->rate_bits;
  p.qdb = qdb;
  p.num_challenges = nc;
  p.num_wires = circ->num_wires;
  p.num_routed = circ->num_routed_wires;
  p.num_constants = circ->num_constants;
  p.num_selectors = circ->num_selectors;
  p.max_degree = qdf;
  p.num_partial_products = (circ->num_routed_wires + qdf - 1) / qdf - 1;  // partial_products.rs:40-47
  if (circ->num_routed_wires == 0) p.num_partial_products = 0;
  p.num_gates = circ->num_gates;
  u32 ngc = 0;
  std::vector<quotient::GateDesc> gates(circ->num_gates);
  for (u32 i = 0; i < circ->num_gates; i++) {
    const p2b_gate& g = circ->gates[i];
    if (g.type >= quotient::G_NUM_TYPES) return fail(P2B_ERR_UNSUPPORTED, "gate %u: unknown gate type %u", i, g.type);
    if (g.selector_index >= circ->num_selectors || g.group_start > i || g.group_end <= i || g.group_end > circ->num_gates)
      return fail(P2B_ERR_INVALID, "gate %u: bad selector group", i);
    if (g.type == quotient::G_RANDOM_ACCESS && g.p0 > 6) return fail(P2B_ERR_UNSUPPORTED, "RandomAccessGate with more than 6 bits");
    if (g.type == quotient::G_COMPARISON && g.p1 == 0) return fail(P2B_ERR_INVALID, "ComparisonGate with zero chunks");
    if ((g.type == quotient::G_HIGH_DEGREE_INTERPOLATION || g.type == quotient::G_LOW_DEGREE_INTERPOLATION) && (g.p0 == 0 || g.p0 > 8))
      return fail(P2B_ERR_UNSUPPORTED, "interpolation gate with subgroup_bits %u outside [1, 8]", g.p0);
    if ((g.type == quotient::G_REDUCING || g.type == quotient::G_REDUCING_EXT || g.type == quotient::G_EXPONENTIATION) && g.p0 == 0)
      return fail(P2B_ERR_INVALID, "gate %u: zero coefficients / power bits", i);
    gates[i] = quotient::GateDesc{g.type, g.selector_index, g.group_start, g.group_end, g.p0, g.p1, g.p2, 0};
    u32 k = quotient::gate_num_constraints(gates[i]);
    ngc = k > ngc ? k : ngc;
  }
  p.num_gate_constraints = ngc;
  p.num_terms = nc * (p.num_partial_products + 2) + ngc;
  for (int i = 0; i < 4; i++) p.pih[i] = pih[i] % gl::P;
  for (u32 i = 0; i < nc; i++) {
    p.betas[i] = betas[i] % gl::P;
    p.gammas[i] = gammas[i] % gl::P;
  }
  // ZeroPolyOnCoset::new (zero_poly_coset.rs:20-33)
  {
    u64 g_pow_n = hostf::pow(hostf::COSET_SHIFT, (u64)1 << circ->degree_bits);
    u64 v = hostf::root(qdb), cur = 1;
    for (u32 i = 0; i < (1u << qdb); i++) {
      u64 gv = hostf::mul(g_pow_n, cur);
      p.zh[i] = gv ? gv - 1 : gl::P - 1;  // g^n * v^i - 1 (no u64 overflow)
      p.zh_inv[i] = hostf::inv(p.zh[i]);
      cur = hostf::mul(cur, v);
    }
  }
  p.w = hostf::root(circ->degree_bits + qdb);
  p.n_field = ((u64)1 << circ->degree_bits) % gl::P;
  for (u32 i = 0; i <= 8; i++) p.small_roots[i] = hostf::root(i);
  p.wires = d_wires;
  p.wires_stride = wires_stride;
  p.zs_pp = d_zs_pp;
  p.zs_stride = zs_stride;
  p.cs = d_cs;
  p.cs_stride = cs_stride;

  const u32 lde_log = circ->degree_bits + qdb;
  const u64 lde_size = (u64)1 << lde_log;
  quotient::GateDesc* d_gates = nullptr;
  u64 *d_kis = nullptr, *d_alphas = nullptr, *d_apows = nullptr, *d_vals = nullptr;
  u32* d_work = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&d_gates, (gates.size() ? gates.size() : 1) * sizeof(quotient::GateDesc), st));
    if (!gates.empty()) CUDA_TRY(cudaMemcpyAsync(d_gates, gates.data(), gates.size() * sizeof(quotient::GateDesc), cudaMemcpyHostToDevice, st));
    CUDA_TRY(pool_alloc(&d_kis, (p.num_routed ? p.num_routed : 1) * sizeof(u64), st));
    if (p.num_routed) CUDA_TRY(cudaMemcpyAsync(d_kis, circ->k_is, p.num_routed * sizeof(u64), cudaMemcpyHostToDevice, st));
    CUDA_TRY(pool_alloc(&d_alphas, nc * sizeof(u64), st));
    u64 ha[quotient::MAX_CHALLENGES];
    for (u32 i = 0; i < nc; i++) ha[i] = alphas[i] % gl::P;
    CUDA_TRY(cudaMemcpyAsync(d_alphas, ha, nc * sizeof(u64), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // host staging buffers (gates, ha) go out of scope
    CUDA_TRY(pool_alloc(&d_apows, (u64)nc * p.num_terms * sizeof(u64), st));
    quotient::alpha_pows_kernel<<<(p.num_terms + 127) / 128, 128, 0, st>>>(d_apows, p.num_terms, nc, d_alphas);
    c->launches++;
    u64* vals = d_values_out;
    if (!vals) {
      CUDA_TRY(pool_alloc(&d_vals, nc * lde_size * sizeof(u64), st));
      vals = d_vals;
    }
    p.pt_first = pt_first;
    p.pt_stride = pt_stride;
    p.pt_count = lde_size / pt_stride;
    p.last_row = last_row == ~(u64)0 ? (((u64)1 << (circ->degree_bits + circ->rate_bits)) - 1) : last_row;
    p.gates = d_gates;
    p.k_is = d_kis;
    p.alpha_pows = d_apows;
    p.out_values = vals;
    p.out_rows = d_rows_out;
    {
      // tile geometry: the most point warps per CTA whose staged rows fit the opt-in shared memory
      quotient::TileGeom tg{};
      tg.nw = p.num_wires;
      tg.ncs = p.num_constants + p.num_routed;
      tg.nzs = nc * (1 + p.num_partial_products);
      tg.ws = quotient::TileGeom::odd_stride(tg.nw);
      tg.css = quotient::TileGeom::odd_stride(tg.ncs);
      tg.zss = quotient::TileGeom::odd_stride(tg.nzs);
      tg.tw = 0;
      for (u32 tw : {6u, 4u, 3u, 2u, 1u}) {
        tg.tw = tw;
        if (tg.words() * sizeof(u64) + 3 * 192 * sizeof(void*) + 1024 <= c->smem_optin) break;
        tg.tw = 0;
      }
      if (!tg.tw) return fail(P2B_ERR_UNSUPPORTED, "circuit rows (%u + %u + %u words) do not fit shared memory", tg.nw, tg.ncs, tg.nzs);
      // work items (every non-trivial gate + the permutation argument), longest first by the instruction-count model of
      // quotient::work_item_cost: the kernel's warp groups pull them in this order
      std::vector<std::pair<u64, u32>> items;
      items.emplace_back(quotient::work_item_cost(p, nullptr), quotient::WORK_PERMUTATION);
      for (u32 i = 0; i < gates.size(); i++)
        if (gates[i].type != quotient::G_NOOP) items.emplace_back(quotient::work_item_cost(p, &gates[i]), i);
      std::stable_sort(items.begin(), items.end(), [](const std::pair<u64, u32>& a, const std::pair<u64, u32>& b2) { return a.first > b2.first; });
      std::vector<u32> flat;
      for (auto& it : items) flat.push_back(it.second);
      CUDA_TRY(pool_alloc(&d_work, flat.size() * sizeof(u32), st));
      CUDA_TRY(cudaMemcpyAsync(d_work, flat.data(), flat.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaStreamSynchronize(st));   // flat goes out of scope
      const u32 num_items = (u32)flat.size();
      const unsigned blocks = (unsigned)((p.pt_count + tg.points() - 1) / tg.points());
      const size_t smem = tg.words() * sizeof(u64);
      constexpr unsigned BD = quotient::QUOT_WARPS * 32;
      switch (nc) {
        case 1:
          P2B_TRY(opt_in_smem(quotient::quotient_values_kernel<1>, smem));
          quotient::quotient_values_kernel<1><<<blocks, BD, smem, st>>>(p, tg, d_work, num_items);
          break;
        case 2:
          P2B_TRY(opt_in_smem(quotient::quotient_values_kernel<2>, smem));
          quotient::quotient_values_kernel<2><<<blocks, BD, smem, st>>>(p, tg, d_work, num_items);
          break;
        case 3:
          P2B_TRY(opt_in_smem(quotient::quotient_values_kernel<3>, smem));
          quotient::quotient_values_kernel<3><<<blocks, BD, smem, st>>>(p, tg, d_work, num_items);
          break;
        default:
          P2B_TRY(opt_in_smem(quotient::quotient_values_kernel<4>, smem));
          quotient::quotient_values_kernel<4><<<blocks, BD, smem, st>>>(p, tg, d_work, num_items);
          break;
      }
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if (d_coeffs_out) P2B_TRY(quotient_values_to_coeffs(c, vals, d_coeffs_out, lde_log, nc));
    return P2B_OK;
  };
  int rc = body();
  for (void* ptr : {(void*)d_gates, (void*)d_kis, (void*)d_alphas, (void*)d_apows, (void*)d_vals, (void*)d_work})
    if (ptr) cudaFreeAsync(ptr, st);
  return rc;
}

extern "C" int p2b_quotient_polys_rows(p2b_ctx* ctx, const p2b_circuit* circuit, const uint64_t* d_wires_rows, uint64_t wires_stride,
                                       const uint64_t* d_zs_pp_rows, uint64_t zs_pp_stride, const uint64_t* d_consts_sigmas_rows,
                                       uint64_t consts_sigmas_stride, const uint64_t* public_inputs_hash, const uint64_t* betas,
                                       const uint64_t* gammas, const uint64_t* alphas, uint64_t* d_values_out, uint64_t* d_coeffs_out) {
  return quotient_impl(ctx, circuit, d_wires_rows, wires_stride, d_zs_pp_rows, zs_pp_stride, d_consts_sigmas_rows, consts_sigmas_stride,
                       public_inputs_hash, betas, gammas, alphas, d_values_out, d_coeffs_out, nullptr);
}

extern "C" int p2b_quotient_polys(p2b_ctx* ctx, const p2b_circuit* circuit, const p2b_batch* wires, const p2b_batch* zs_pp,
                                  const p2b_batch* consts_sigmas, const uint64_t* public_inputs_hash, const uint64_t* betas,
                                  const uint64_t* gammas, const uint64_t* alphas, uint64_t* d_values_out, uint64_t* d_coeffs_out) {
  if (!circuit || !wires || !zs_pp || !consts_sigmas) return fail(P2B_ERR_INVALID, "NULL argument");
  for (const p2b_batch* b : {wires, zs_pp, consts_sigmas}) {
    if (b->info.degree_log != circuit->degree_bits || b->info.rate_bits != circuit->rate_bits)
      return fail(P2B_ERR_INVALID, "batch shape does not match the circuit (degree_bits / rate_bits)");
    if (b->local_leaves != b->info.num_leaves) return fail(P2B_ERR_UNSUPPORTED, "sharded batches: gather the rows first");
  }
  u32 qdf = circuit->quotient_degree_factor;
  u64 npp = circuit->num_routed_wires ? (circuit->num_routed_wires + qdf - 1) / (qdf ? qdf : 1) - 1 : 0;
  if (wires->info.num_polys < circuit->num_wires) return fail(P2B_ERR_INVALID, "wires batch has fewer columns than num_wires");
  if (consts_sigmas->info.num_polys < (u64)circuit->num_constants + circuit->num_routed_wires)
    return fail(P2B_ERR_INVALID, "constants/sigmas batch has too few columns");
  if (zs_pp->info.num_polys < (u64)circuit->num_challenges * (1 + npp)) return fail(P2B_ERR_INVALID, "zs/partial-products batch has too few columns");
  return quotient_impl(ctx, circuit, wires->leaves, wires->info.leaf_len, zs_pp->leaves, zs_pp->info.leaf_len, consts_sigmas->leaves,
                       consts_sigmas->info.leaf_len, public_inputs_hash, betas, gammas, alphas, d_values_out, d_coeffs_out, nullptr);
}


// ======================================================================================================
// permutation argument: Z and partial products (plonk/prover.rs:702-786, :112-117)
// ======================================================================================================
extern "C" int p2b_partial_products_and_zs(p2b_ctx* c, const uint64_t* d_wires_values, const uint64_t* d_sigma_values,
                                           uint32_t degree_bits, uint32_t num_routed_wires, uint32_t quotient_degree_factor,
                                           uint32_t num_challenges, const uint64_t* k_is, const uint64_t* betas,
                                           const uint64_t* gammas, uint64_t* d_out) {
  if (!c || !d_wires_values || !d_sigma_values || !k_is || !betas || !gammas || !d_out) return fail(P2B_ERR_INVALID, "NULL argument");
  Stage stage(c, c->stream, "compute partial products");   // plonk/prover.rs:112-117
  if (degree_bits > 32) return fail(P2B_ERR_INVALID, "degree_bits exceeds the field's two-adicity 32");
  if (num_challenges == 0 || num_challenges > (u32)perm::MAX_CH) return fail(P2B_ERR_INVALID, "num_challenges must be in [1, %d]", perm::MAX_CH);
  if (quotient_degree_factor < 2) return fail(P2B_ERR_INVALID, "quotient_degree_factor must be at least 2");
  if (!(quotient_degree_factor < num_routed_wires))  // prover.rs:102-105
    return fail(P2B_ERR_INVALID, "When the number of routed wires is smaller that the degree, we should change the logic to avoid computing partial products.");
  const u32 K = (num_routed_wires + quotient_degree_factor - 1) / quotient_degree_factor;
  if (K > 64) return fail(P2B_ERR_UNSUPPORTED, "more than 64 partial-product chunks (%u)", K);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const u64 n = (u64)1 << degree_bits;
  const u32 nb = (u32)((n + perm::SCAN_BLOCK - 1) / perm::SCAN_BLOCK);
  perm::Challenges ch{};
  for (u32 i = 0; i < num_challenges; i++) {
    ch.beta[i] = betas[i] % gl::P;
    ch.gamma[i] = gammas[i] % gl::P;
  }
  u64 *d_kis = nullptr, *d_tot = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(pool_alloc(&d_kis, num_routed_wires * sizeof(u64), st));
    CUDA_TRY(pool_alloc(&d_tot, ((u64)num_challenges * nb + 1) * sizeof(u64), st));
    u32* d_flag = reinterpret_cast<u32*>(d_tot + (u64)num_challenges * nb);
    CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(u64), st));
    CUDA_TRY(cudaMemcpyAsync(d_kis, k_is, num_routed_wires * sizeof(u64), cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + 127) / 128);
    const u64 w = hostf::root(degree_bits);
    if (K <= 16)
      perm::chunk_products_kernel<16><<<blocks, 128, 0, st>>>(d_wires_values, d_sigma_values, n, degree_bits, num_routed_wires,
                                                              quotient_degree_factor, num_challenges, ch, d_kis, w, d_out, d_flag);
    else
      perm::chunk_products_kernel<64><<<blocks, 128, 0, st>>>(d_wires_values, d_sigma_values, n, degree_bits, num_routed_wires,
                                                              quotient_degree_factor, num_challenges, ch, d_kis, w, d_out, d_flag);
    perm::block_totals_kernel<<<dim3(nb, num_challenges), perm::SCAN_BLOCK, 0, st>>>(d_out, n, nb, d_tot);
    perm::scan_totals_kernel<<<num_challenges, perm::SCAN_BLOCK, 0, st>>>(d_tot, nb);
    perm::apply_kernel<<<dim3(nb, num_challenges), perm::SCAN_BLOCK, 0, st>>>(d_out, n, nb, num_challenges, K, d_tot);
    c->launches += 4;
    CUDA_TRY(cudaGetLastError());
    // also covers the host k_is staging above (pinned callers may reuse the buffer once this returns)
    u32 flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (flag) return fail(P2B_ERR_INVALID, "Tried to invert zero (a permutation denominator vanished; field/src/types.rs:130)");
    return P2B_OK;
  };
  int rc = body();
  if (d_kis) cudaFreeAsync(d_kis, st);
  if (d_tot) cudaFreeAsync(d_tot, st);
  return rc;
}

#include "fri_api.cuh"
#include "compat.cuh"
#include "mgpu.cuh"
