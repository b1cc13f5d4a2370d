// mgpu.cuh -- single-process, multi-device commit: the shape the reference's caller has (ONE prover process,
// plonky2/src/fri/oracle.rs:279-545 `from_values_with_gpu`, plonky2/src/plonk/prover.rs:239-700 `my_prove`), so a Rust
// prover can drive all GPUs of a node through the C ABI.  Included at the end of plonky2_b200.cu (single translation unit).
//
// Partition (the same as plonky2-gpu_b200/sharded.py, which drives one process per GPU for bench.py):
//   * the value matrix is dealt in exchange rounds of 8, 8, 16, 32, .. 8G CONSECUTIVE columns (mgpu_schedule) -- the order
//     in which the leaves' sponges absorb them (hashing.rs:81-104) -- and device s owns `per` consecutive columns of each
//     round; the first rounds are small so hashing starts after one column per device has been transformed and exchanged;
//   * device g uploads and inverse-transforms its blocks, then PUSHES each block into every device's coefficient matrix
//     with cudaMemcpyPeerAsync over NVLink on its own copy stream (the one real exchange of the path; no NCCL needed in
//     one process);
//   * device d owns coset blocks [d R/G, (d+1) R/G) = the contiguous leaf range [d N/G, (d+1) N/G): as soon as a round has
//     landed (cross-device cudaStreamWaitEvent -- no host synchronisation anywhere in the flow) it low-degree-extends those
//     columns into its leaf rows and advances its leaves' sponges while the next round is in flight;
//   * digest layers up to the device's top layer, then the top-layer nodes (the cap entries whenever 2^cap_height >= G)
//     are pushed to every device and the remaining layers, if any, are finished everywhere.
// The host thread only enqueues; p2b_mgpu_synchronize() / the getters wait.
#pragma once

struct p2b_mgpu {
  int n = 0;
  std::vector<p2b_ctx*> ctx;
  std::vector<cudaStream_t> xfer;                 // per device: the stream its peer pushes run on
  std::vector<std::vector<cudaEvent_t>> ev_ifft;  // [device][round]
  std::vector<std::vector<cudaEvent_t>> ev_push;  // [device][round]
  std::vector<cudaEvent_t> ev_nodes, ev_t0, ev_t1, ev_done;
  std::vector<std::vector<cudaEvent_t>> ev_fwd;   // [device]: events of that device's copy stream (ONE_DEVICE source forwarding)
  std::vector<u64*> cols;                         // per device: its column blocks [rounds * 8][n]
  std::vector<u64> cols_elems;
  std::vector<u64*> nodes_all;                    // per device: gathered top-layer nodes
  std::vector<u64> nodes_elems;
  bool peer_ok = true;
  bool single_thread_issue = false;               // P2B_MGPU_SINGLE_THREAD=1: enqueue for every device from the caller's thread (A/B, debugging)
};

struct p2b_mgpu_batch {
  p2b_mgpu* g = nullptr;
  std::vector<p2b_batch*> shard;  // one per device, leaves [d N/G, (d+1) N/G)
  p2b_batch_info info{};
};

// exchange rounds (mirrors sharded.exchange_schedule / local_layout in plonky2-gpu_b200/sharded.py)
struct MgpuRound {
  u64 col0, width, per, row0;  // columns [col0, col0 + width); device s owns [col0 + s*per, +per) clipped, at local rows [row0, row0 + per)
};
static std::vector<MgpuRound> mgpu_schedule(u64 P, int G, u64* rows_total) {
  std::vector<MgpuRound> r;
  u64 col0 = 0, w = 8, row0 = 0;
  while (col0 < P) {
    u64 width = std::min<u64>(w, P - col0), per = (width + G - 1) / G;
    r.push_back(MgpuRound{col0, width, per, row0});
    col0 += width;
    row0 += per;
    if (r.size() >= 2) w = std::min<u64>(2 * w, (u64)8 * G);
  }
  if (rows_total) *rows_total = row0;
  return r;
}

static int mgpu_events(p2b_mgpu* g, int dev, size_t rounds) {
  CUDA_TRY(cudaSetDevice(g->ctx[dev]->device));
  while (g->ev_ifft[dev].size() < rounds) {
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    g->ev_ifft[dev].push_back(a);
    g->ev_push[dev].push_back(b);
  }
  return P2B_OK;
}

extern "C" void p2b_mgpu_destroy(p2b_mgpu* g) {
  if (!g) return;
  for (int d = 0; d < g->n; d++) {
    if (!g->ctx[d]) continue;
    cudaSetDevice(g->ctx[d]->device);
    if (g->xfer[d]) {
      cudaStreamSynchronize(g->xfer[d]);
      cudaStreamDestroy(g->xfer[d]);
    }
    cudaStreamSynchronize(g->ctx[d]->stream);
    for (cudaEvent_t e : g->ev_ifft[d]) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ev_push[d]) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ev_fwd[d]) cudaEventDestroy(e);
    for (cudaEvent_t e : {g->ev_nodes[d], g->ev_t0[d], g->ev_t1[d], g->ev_done[d]})
      if (e) cudaEventDestroy(e);
    if (g->cols[d]) cudaFree(g->cols[d]);
    if (g->nodes_all[d]) cudaFree(g->nodes_all[d]);
    p2b_ctx_destroy(g->ctx[d]);
  }
  delete g;
}

extern "C" int p2b_mgpu_create(const int* devices, int n_dev, p2b_mgpu** out) {
  if (!out) return fail(P2B_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_dev <= 0 || (n_dev & (n_dev - 1)) || n_dev > 64) return fail(P2B_ERR_INVALID, "device count %d must be a power of two in [1, 64]", n_dev);
  p2b_mgpu* g = new (std::nothrow) p2b_mgpu();
  if (!g) return fail(P2B_ERR_OOM, "host allocation failed");
  g->n = n_dev;
  g->single_thread_issue = getenv("P2B_MGPU_SINGLE_THREAD") != nullptr;
  g->ctx.assign(n_dev, nullptr);
  g->xfer.assign(n_dev, nullptr);
  g->ev_ifft.resize(n_dev);
  g->ev_push.resize(n_dev);
  g->ev_fwd.resize(n_dev);
  g->ev_nodes.assign(n_dev, nullptr);
  g->ev_t0.assign(n_dev, nullptr);
  g->ev_t1.assign(n_dev, nullptr);
  g->ev_done.assign(n_dev, nullptr);
  g->cols.assign(n_dev, nullptr);
  g->cols_elems.assign(n_dev, 0);
  g->nodes_all.assign(n_dev, nullptr);
  g->nodes_elems.assign(n_dev, 0);
  auto body = [&]() -> int {
    for (int d = 0; d < n_dev; d++) {
      int dev = devices ? devices[d] : d;
      for (int e = 0; e < d; e++)
        if (g->ctx[e]->device == dev) return fail(P2B_ERR_INVALID, "device %d listed twice", dev);
      P2B_TRY(p2b_ctx_create(dev, &g->ctx[d]));
      CUDA_TRY(cudaStreamCreateWithFlags(&g->xfer[d], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&g->ev_nodes[d], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&g->ev_done[d], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreate(&g->ev_t0[d]));
      CUDA_TRY(cudaEventCreate(&g->ev_t1[d]));
    }
    // direct NVLink copies between every pair; the stream-ordered pools (batch arrays) must be mapped into the peers too
    for (int a = 0; a < n_dev; a++) {
      CUDA_TRY(cudaSetDevice(g->ctx[a]->device));
      cudaMemPool_t pool;
      CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, g->ctx[a]->device));
      for (int b = 0; b < n_dev; b++) {
        if (a == b) continue;
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, g->ctx[a]->device, g->ctx[b]->device));
        if (!can) {
          g->peer_ok = false;  // copies still work (staged by the driver), only slower
          continue;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[b]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(P2B_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaMemAccessDesc desc{};
        desc.location.type = cudaMemLocationTypeDevice;
        desc.location.id = g->ctx[b]->device;
        desc.flags = cudaMemAccessFlagsProtReadWrite;
        CUDA_TRY(cudaMemPoolSetAccess(pool, &desc, 1));
      }
    }
    return P2B_OK;
  };
  int rc = body();
  if (rc != P2B_OK) {
    p2b_mgpu_destroy(g);
    return rc;
  }
  *out = g;
  return P2B_OK;
}

extern "C" int p2b_mgpu_device_count(const p2b_mgpu* g) { return g ? g->n : 0; }
extern "C" p2b_ctx* p2b_mgpu_ctx(p2b_mgpu* g, int index) { return (g && index >= 0 && index < g->n) ? g->ctx[index] : nullptr; }
extern "C" int p2b_mgpu_peer_access(const p2b_mgpu* g) { return g && g->peer_ok ? 1 : 0; }

extern "C" int p2b_mgpu_synchronize(p2b_mgpu* g) {
  if (!g) return fail(P2B_ERR_INVALID, "NULL argument");
  for (int d = 0; d < g->n; d++) {
    CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
    CUDA_TRY(cudaStreamSynchronize(g->xfer[d]));
    P2B_TRY(p2b_ctx_synchronize(g->ctx[d]));
  }
  return P2B_OK;
}

// device-side stopwatch over all devices: start marks every main stream, stop returns the longest span in ms
extern "C" int p2b_mgpu_timer_start(p2b_mgpu* g) {
  if (!g) return fail(P2B_ERR_INVALID, "NULL argument");
  for (int d = 0; d < g->n; d++) {
    CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
    CUDA_TRY(cudaEventRecord(g->ev_t0[d], g->ctx[d]->stream));
  }
  return P2B_OK;
}
extern "C" int p2b_mgpu_timer_stop_ms(p2b_mgpu* g, float* ms_max) {
  if (!g || !ms_max) return fail(P2B_ERR_INVALID, "NULL argument");
  *ms_max = 0;
  for (int d = 0; d < g->n; d++) {
    CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
    CUDA_TRY(cudaEventRecord(g->ev_t1[d], g->ctx[d]->stream));
  }
  for (int d = 0; d < g->n; d++) {
    CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
    CUDA_TRY(cudaEventSynchronize(g->ev_t1[d]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev_t0[d], g->ev_t1[d]));
    if (ms > *ms_max) *ms_max = ms;
  }
  return P2B_OK;
}

// every device pushes its `count` top-layer nodes (layer `top`) to all peers, imports everyone's and finishes the layers above
static int mgpu_exchange_nodes(p2b_mgpu* g, p2b_mgpu_batch* mb, u32 top, u64 count) {
  const int G = g->n;
  for (int d = 0; d < G; d++) {
    CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
    P2B_TRY(mgpu_events(g, d, 1));
    if (g->nodes_elems[d] < (u64)G * count * 4) {
      if (g->nodes_all[d]) {
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaFree(g->nodes_all[d]));
      }
      CUDA_TRY(cudaMalloc(&g->nodes_all[d], (u64)G * count * 4 * sizeof(u64)));
      g->nodes_elems[d] = (u64)G * count * 4;
    }
  }
  for (int s = 0; s < G; s++) {
    p2b_ctx* c = g->ctx[s];
    CUDA_TRY(cudaSetDevice(c->device));
    u64* mine = g->nodes_all[s] + (u64)s * count * 4;
    P2B_TRY(p2b_batch_export_nodes(mb->shard[s], top, (u64)s * count, count, mine));
    CUDA_TRY(cudaEventRecord(g->ev_ifft[s][0], c->stream));
    CUDA_TRY(cudaStreamWaitEvent(g->xfer[s], g->ev_ifft[s][0], 0));
    for (int d = 0; d < G; d++)
      if (d != s)
        CUDA_TRY(cudaMemcpyPeerAsync(g->nodes_all[d] + (u64)s * count * 4, g->ctx[d]->device, mine, c->device, count * 4 * sizeof(u64), g->xfer[s]));
    CUDA_TRY(cudaEventRecord(g->ev_nodes[s], g->xfer[s]));
  }
  for (int d = 0; d < G; d++) {
    p2b_ctx* c = g->ctx[d];
    CUDA_TRY(cudaSetDevice(c->device));
    for (int s = 0; s < G; s++)
      if (s != d) CUDA_TRY(cudaStreamWaitEvent(c->stream, g->ev_nodes[s], 0));
    P2B_TRY(p2b_batch_import_nodes(mb->shard[d], top, 0, (u64)G * count, g->nodes_all[d]));
    P2B_TRY(p2b_batch_finish_layers(mb->shard[d], top));
  }
  return P2B_OK;
}

extern "C" void p2b_mgpu_batch_destroy(p2b_mgpu_batch* b) {
  if (!b) return;
  for (p2b_batch* s : b->shard)
    if (s) batch_free(s);   // checks that the shard's context is still alive
  delete b;
}

// where the caller's value columns live
enum { P2B_MGPU_SRC_HOST = 0, P2B_MGPU_SRC_RESIDENT = 1, P2B_MGPU_SRC_ONE_DEVICE = 2 };

// PolynomialBatch::from_values over all devices of `g` (fri/oracle.rs:709-731 / :279-545).
//   src == HOST    : values = host [P][n] (pinned for overlap); uploaded block by block behind the transforms
//   src == RESIDENT: the column blocks are already on their devices in the layout of p2b_mgpu_resident_cols()
//                    (what a prover that generates its witness on the GPUs, or a benchmark with inputs in HBM, has)
//   coeffs_host_out: NULL or host [P][n]; receives the coefficients (the reference keeps them host-side, oracle.rs:403-407)
//   src == ONE_DEVICE: values = device pointer [P][n] on device index `src_index` (e.g. the Z / partial-product matrix the
//                    first device just computed): its columns are pushed to their owners over NVLink first
static int mgpu_commit(p2b_mgpu* g, int src, const u64* values, u32 k, u64 P, u32 rate_bits, u32 cap_height, u64* coeffs_host_out,
                       p2b_mgpu_batch** out, int src_index = 0) {
  if (!g || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = nullptr;
  const int G = g->n;
  const u64 R = (u64)1 << rate_bits;
  if ((u64)G > R) return fail(P2B_ERR_INVALID, "%d devices exceed the 2^rate_bits = %llu coset blocks", G, (unsigned long long)R);
  if (P <= 4) return fail(P2B_ERR_UNSUPPORTED, "multi-device commit needs more than 4 polynomials (hash_or_noop copies shorter leaves)");
  if (src != P2B_MGPU_SRC_RESIDENT && !values) return fail(P2B_ERR_INVALID, "NULL values");
  if (src == P2B_MGPU_SRC_ONE_DEVICE && (src_index < 0 || src_index >= G)) return fail(P2B_ERR_INVALID, "source device index out of range");
  const u64 n = (u64)1 << k;
  u64 rows_total = 0;
  const std::vector<MgpuRound> sched = mgpu_schedule(P, G, &rows_total);
  const u64 rounds = sched.size();
  p2b_mgpu_batch* mb = new (std::nothrow) p2b_mgpu_batch();
  if (!mb) return fail(P2B_ERR_OOM, "host allocation failed");
  mb->g = g;
  mb->shard.assign(G, nullptr);
  auto body = [&]() -> int {
    for (int d = 0; d < G; d++) {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      P2B_TRY(mgpu_events(g, d, (size_t)rounds));
      if (g->cols_elems[d] < rows_total * n) {
        if (src == P2B_MGPU_SRC_RESIDENT) return fail(P2B_ERR_INVALID, "resident columns were not allocated for this shape (p2b_mgpu_resident_cols)");
        if (g->cols[d]) {
          CUDA_TRY(cudaDeviceSynchronize());
          CUDA_TRY(cudaFree(g->cols[d]));
          g->cols[d] = nullptr;
        }
        CUDA_TRY(cudaMalloc(&g->cols[d], rows_total * n * sizeof(u64)));
        g->cols_elems[d] = rows_total * n;
      }
      P2B_TRY(ensure_scratch(c, (u64)G * 8 * n));
      P2B_TRY(p2b_commit_blocks_begin(c, k, P, rate_bits, cap_height, (u64)d * (R / G), R / G, &mb->shard[d]));
      // the peers push into this shard's coefficient matrix: its allocation must precede their copies; and this device's
      // column buffer may still be read by an earlier commit: uploads into it are ordered behind the main stream as well
      CUDA_TRY(cudaEventRecord(g->ev_done[d], c->stream));
      CUDA_TRY(cudaStreamWaitEvent(c->stream_h2d, g->ev_done[d], 0));
    }
    for (int s = 0; s < G; s++)
      for (int d = 0; d < G; d++) {
        CUDA_TRY(cudaSetDevice(g->ctx[s]->device));
        CUDA_TRY(cudaStreamWaitEvent(g->xfer[s], g->ev_done[d], 0));
      }
    auto slice = [&](const MgpuRound& r, int s, u64* c0, u64* nc) {
      *c0 = std::min<u64>(r.col0 + (u64)s * r.per, r.col0 + r.width);
      *nc = std::min<u64>(r.per, r.col0 + r.width - *c0);
    };
    // stage B of (device d, round j): once every slice of round j has landed, low-degree-extend those columns into d's leaf
    // rows and advance its sponges.  `published` (threaded issue only): the host-side record that device s has RECORDED
    // ev_push[s][j] -- cudaStreamWaitEvent on an event that has not been recorded yet would not wait at all.
    std::vector<std::atomic<u64>> published(G);
    for (auto& a : published) a.store(0, std::memory_order_relaxed);
    std::atomic<int> abort_flag{0};
    auto absorb_one = [&](int d, u64 j, bool threaded) -> int {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      for (int s = 0; s < G; s++) {
        u64 c0, nc;
        slice(sched[j], s, &c0, &nc);
        if (!nc) continue;
        if (threaded)
          while (published[s].load(std::memory_order_acquire) <= j) {
            if (abort_flag.load(std::memory_order_relaxed)) return fail(P2B_ERR_CUDA, "multi-device commit aborted: another device's issue thread failed");
            std::this_thread::yield();
          }
        CUDA_TRY(cudaStreamWaitEvent(c->stream, g->ev_push[s][j], 0));
      }
      return lde_and_absorb_group(mb->shard[d], sched[j].col0, sched[j].width, true);
    };
    // stage A of (device s, round j): upload (or receive) its slice, inverse-transform it, copy the coefficients out and push
    // them to every device
    auto produce_one = [&](int s, u64 j) -> int {
      u64 c0, nc;
      slice(sched[j], s, &c0, &nc);
      if (!nc) return P2B_OK;
      p2b_ctx* c = g->ctx[s];
      CUDA_TRY(cudaSetDevice(c->device));
      u64* blk = g->cols[s] + sched[j].row0 * n;
      if (src == P2B_MGPU_SRC_HOST) {
        cudaEvent_t up = c->ev_copy[j % 14];
        CUDA_TRY(cudaMemcpyAsync(blk, values + c0 * n, nc * n * sizeof(u64), cudaMemcpyHostToDevice, c->stream_h2d));
        CUDA_TRY(cudaEventRecord(up, c->stream_h2d));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, up, 0));
      } else if (src == P2B_MGPU_SRC_ONE_DEVICE) {
        // the source device's main stream produced `values`; its copy stream forwards this slice to its owner
        p2b_ctx* sc = g->ctx[src_index];
        if (j == 0 && s == 0) {
          CUDA_TRY(cudaSetDevice(sc->device));
          CUDA_TRY(cudaEventRecord(sc->ev_a, sc->stream));
          CUDA_TRY(cudaStreamWaitEvent(g->xfer[src_index], sc->ev_a, 0));
        }
        CUDA_TRY(cudaSetDevice(sc->device));
        CUDA_TRY(cudaMemcpyPeerAsync(blk, c->device, values + c0 * n, sc->device, nc * n * sizeof(u64), g->xfer[src_index]));
        const size_t slot = (size_t)j * G + s;   // an event of the SOURCE device (events are recorded on their own device's streams)
        while (g->ev_fwd[src_index].size() <= slot) {
          cudaEvent_t e;
          CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          g->ev_fwd[src_index].push_back(e);
        }
        cudaEvent_t up = g->ev_fwd[src_index][slot];
        CUDA_TRY(cudaEventRecord(up, g->xfer[src_index]));
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, up, 0));
      }
      P2B_TRY(run_ifft(c, blk, blk, c->scratch, k, nc));
      CUDA_TRY(cudaEventRecord(g->ev_ifft[s][j], c->stream));
      if (coeffs_host_out) {
        CUDA_TRY(cudaStreamWaitEvent(c->stream_d2h, g->ev_ifft[s][j], 0));
        CUDA_TRY(cudaMemcpyAsync(coeffs_host_out + c0 * n, blk, nc * n * sizeof(u64), cudaMemcpyDeviceToHost, c->stream_d2h));
      }
      CUDA_TRY(cudaStreamWaitEvent(g->xfer[s], g->ev_ifft[s][j], 0));
      for (int dd = 0; dd < G; dd++) {
        const int d = (s + dd) % G;  // stagger the destinations so the pushes of one round spread over the switch
        CUDA_TRY(cudaMemcpyPeerAsync(mb->shard[d]->coeffs + c0 * n, g->ctx[d]->device, blk, c->device, nc * n * sizeof(u64), g->xfer[s]));
      }
      CUDA_TRY(cudaEventRecord(g->ev_push[s][j], g->xfer[s]));
      return P2B_OK;
    };
    // One issue thread per device (host input or resident input, G > 1): a commit is ~60-80 runtime calls per device, and one
    // thread enqueueing for eight devices was the end-to-end limiter (27.6 ms against 18.5 ms of device time at 8 GPUs).  Each
    // thread runs its device's program in round order; the only cross-thread dependency is "wait for an event another thread
    // records", ordered on the host through `published`.  Forwarding from one device keeps the single-thread order (its events
    // live on the source device's copy stream).
    const bool threaded = G > 1 && src != P2B_MGPU_SRC_ONE_DEVICE && !g->single_thread_issue;
    if (threaded) {
      std::vector<int> rcs(G, P2B_OK);
      std::vector<std::string> errs(G);
      auto program = [&](int s) {
        int rc = P2B_OK;
        for (u64 j = 0; j < rounds && rc == P2B_OK; j++) {
          rc = produce_one(s, j);
          if (rc == P2B_OK) published[s].store(j + 1, std::memory_order_release);
          if (rc == P2B_OK && j >= 1) rc = absorb_one(s, j - 1, true);
        }
        if (rc == P2B_OK) rc = absorb_one(s, rounds - 1, true);
        if (rc != P2B_OK) {
          errs[s] = g_last_error;   // thread-local in the worker: handed to the caller's thread below
          abort_flag.store(1, std::memory_order_relaxed);
        }
        rcs[s] = rc;
      };
      std::vector<std::thread> workers;
      bool spawn_failed = false;
      try {
        workers.reserve(G - 1);
        for (int s = 1; s < G; s++) workers.emplace_back(program, s);
      } catch (...) {   // no exception crosses the C ABI: the threads already running are told to stop and joined
        spawn_failed = true;
        abort_flag.store(1, std::memory_order_relaxed);
      }
      if (!spawn_failed) program(0);
      for (auto& w : workers) w.join();
      if (spawn_failed) return fail(P2B_ERR_CUDA, "multi-device commit: could not start the per-device issue threads (set P2B_MGPU_SINGLE_THREAD=1)");
      for (int s = 0; s < G; s++)
        if (rcs[s] != P2B_OK) {
          g_last_error = errs[s];
          return rcs[s];
        }
    } else {
      for (u64 j = 0; j < rounds; j++) {
        for (int s = 0; s < G; s++) P2B_TRY(produce_one(s, j));
        if (j >= 1)
          for (int d = 0; d < G; d++) P2B_TRY(absorb_one(d, j - 1, false));
      }
      for (int d = 0; d < G; d++) P2B_TRY(absorb_one(d, rounds - 1, false));
    }
    // digest layers on every device, then the top-layer node exchange
    u32 top = 0;
    u64 count = 0;
    for (int d = 0; d < G; d++) {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      // lde_and_absorb_group hashed on stream2: join before the layers
      CUDA_TRY(cudaEventRecord(c->ev_b, c->stream2));
      CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
      P2B_TRY(p2b_commit_blocks_finish(mb->shard[d]));
      top = mb->shard[d]->top_layer;
      count = mb->shard[d]->local_leaves >> top;
    }
    if (G > 1) P2B_TRY(mgpu_exchange_nodes(g, mb, top, count));
    if (coeffs_host_out)
      for (int d = 0; d < G; d++) {
        CUDA_TRY(cudaSetDevice(g->ctx[d]->device));
        CUDA_TRY(cudaEventRecord(g->ctx[d]->ev_a, g->ctx[d]->stream_d2h));
        CUDA_TRY(cudaStreamWaitEvent(g->ctx[d]->stream, g->ctx[d]->ev_a, 0));
      }
    return P2B_OK;
  };
  int rc = body();
  if (rc != P2B_OK) {
    for (int d = 0; d < G; d++)
      if (g->ctx[d]) {
        cudaSetDevice(g->ctx[d]->device);
        cudaDeviceSynchronize();
      }
    p2b_mgpu_batch_destroy(mb);
    return rc;
  }
  mb->info = mb->shard[0]->info;
  *out = mb;
  return P2B_OK;
}

extern "C" int p2b_mgpu_commit_from_values(p2b_mgpu* g, const uint64_t* values_host, uint32_t degree_log, uint64_t num_polys,
                                           uint32_t rate_bits, uint32_t cap_height, uint64_t* coeffs_host_out, p2b_mgpu_batch** out) {
  return mgpu_commit(g, P2B_MGPU_SRC_HOST, values_host, degree_log, num_polys, rate_bits, cap_height, coeffs_host_out, out);
}

extern "C" int p2b_mgpu_commit_from_device_values(p2b_mgpu* g, int src_index, const uint64_t* d_values, uint32_t degree_log,
                                                  uint64_t num_polys, uint32_t rate_bits, uint32_t cap_height, p2b_mgpu_batch** out) {
  return mgpu_commit(g, P2B_MGPU_SRC_ONE_DEVICE, d_values, degree_log, num_polys, rate_bits, cap_height, nullptr, out, src_index);
}

// Device-resident input: returns (allocating on first use) device `index`'s local buffer [rows_total][n] in the layout of
// the exchange schedule: for round j (p2b_mgpu_round), rows [row0_j, row0_j + per_j) hold this device's columns
// [col0_j + index * per_j, ...) of the value matrix (zero rows where the round is ragged).
extern "C" int p2b_mgpu_resident_cols(p2b_mgpu* g, int index, uint32_t degree_log, uint64_t num_polys, uint64_t** d_cols_out,
                                      uint64_t* rounds_out) {
  if (!g || index < 0 || index >= g->n || !d_cols_out) return fail(P2B_ERR_INVALID, "bad argument");
  const u64 n = (u64)1 << degree_log;
  u64 rows_total = 0;
  const u64 rounds = mgpu_schedule(num_polys, g->n, &rows_total).size();
  CUDA_TRY(cudaSetDevice(g->ctx[index]->device));
  if (g->cols_elems[index] < rows_total * n) {
    if (g->cols[index]) {
      CUDA_TRY(cudaDeviceSynchronize());
      CUDA_TRY(cudaFree(g->cols[index]));
      g->cols[index] = nullptr;
    }
    CUDA_TRY(cudaMalloc(&g->cols[index], rows_total * n * sizeof(u64)));
    CUDA_TRY(cudaMemset(g->cols[index], 0, rows_total * n * sizeof(u64)));
    g->cols_elems[index] = rows_total * n;
  }
  *d_cols_out = g->cols[index];
  if (rounds_out) *rounds_out = rounds;
  return P2B_OK;
}
// round j of the exchange schedule for (num_polys, this device count): columns [col0, col0 + width), `per` per device, at local
// rows [row0, row0 + per); returns P2B_ERR_INVALID past the last round
extern "C" int p2b_mgpu_round(const p2b_mgpu* g, uint64_t num_polys, uint64_t j, uint64_t* col0, uint64_t* width, uint64_t* per,
                              uint64_t* row0) {
  if (!g) return fail(P2B_ERR_INVALID, "NULL argument");
  const std::vector<MgpuRound> sched = mgpu_schedule(num_polys, g->n, nullptr);
  if (j >= sched.size()) return fail(P2B_ERR_INVALID, "round %llu past the %zu rounds of the schedule", (unsigned long long)j, sched.size());
  if (col0) *col0 = sched[j].col0;
  if (width) *width = sched[j].width;
  if (per) *per = sched[j].per;
  if (row0) *row0 = sched[j].row0;
  return P2B_OK;
}
// NOTE: the transform is done in place -- the resident values are consumed (refill them before the next commit).
extern "C" int p2b_mgpu_commit_resident(p2b_mgpu* g, uint32_t degree_log, uint64_t num_polys, uint32_t rate_bits, uint32_t cap_height,
                                        uint64_t* coeffs_host_out, p2b_mgpu_batch** out) {
  return mgpu_commit(g, P2B_MGPU_SRC_RESIDENT, nullptr, degree_log, num_polys, rate_bits, cap_height, coeffs_host_out, out);
}

extern "C" int p2b_mgpu_batch_get_info(const p2b_mgpu_batch* b, p2b_batch_info* out) {
  if (!b || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = b->info;
  return P2B_OK;
}
extern "C" p2b_batch* p2b_mgpu_batch_shard(p2b_mgpu_batch* b, int index) {
  return (b && index >= 0 && index < (int)b->shard.size()) ? b->shard[index] : nullptr;
}
// the full cap (identical on every device after the exchange); waits for the commit
extern "C" int p2b_mgpu_batch_get_cap(const p2b_mgpu_batch* b, uint64_t* out) {
  if (!b || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  P2B_TRY(p2b_mgpu_synchronize(b->g));
  return p2b_batch_get_cap(b->shard[0], out);
}
// FRI query openings (fri/prover.rs:187-216): every index is served by the device that owns the leaf.
extern "C" int p2b_mgpu_batch_open_rows(const p2b_mgpu_batch* b, const uint64_t* leaf_indices, uint64_t count, uint64_t* rows_out,
                                        uint64_t* siblings_out) {
  if (!b || !leaf_indices || !rows_out) return fail(P2B_ERR_INVALID, "NULL argument");
  const u64 per = b->info.num_leaves / b->shard.size();
  const u64 ll = b->info.leaf_len, layers = b->shard[0]->shape.sub_log;
  P2B_TRY(p2b_mgpu_synchronize(b->g));
  for (size_t d = 0; d < b->shard.size(); d++) {
    std::vector<u64> idx;
    std::vector<u64> slot;
    for (u64 i = 0; i < count; i++) {
      if (leaf_indices[i] >= b->info.num_leaves) return fail(P2B_ERR_INVALID, "leaf index %llu out of range", (unsigned long long)leaf_indices[i]);
      if (leaf_indices[i] / per == d) {
        idx.push_back(leaf_indices[i]);
        slot.push_back(i);
      }
    }
    if (idx.empty()) continue;
    std::vector<u64> rows(idx.size() * ll), sibs(siblings_out ? idx.size() * layers * 4 : 0);
    P2B_TRY(p2b_batch_open_rows(b->shard[d], idx.data(), idx.size(), rows.data(), siblings_out && layers ? sibs.data() : nullptr));
    for (size_t t = 0; t < idx.size(); t++) {
      memcpy(rows_out + slot[t] * ll, rows.data() + t * ll, ll * sizeof(u64));
      if (siblings_out && layers) memcpy(siblings_out + slot[t] * layers * 4, sibs.data() + t * layers * 4, layers * 4 * sizeof(u64));
    }
  }
  return P2B_OK;
}
// rows [first_leaf, first_leaf + count) of the LDE matrix, across shard boundaries
extern "C" int p2b_mgpu_batch_get_leaves(const p2b_mgpu_batch* b, uint64_t first_leaf, uint64_t count, uint64_t* out) {
  if (!b || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  if (first_leaf + count > b->info.num_leaves) return fail(P2B_ERR_INVALID, "leaf range out of bounds");
  P2B_TRY(p2b_mgpu_synchronize(b->g));
  const u64 per = b->info.num_leaves / b->shard.size();
  u64 done = 0;
  while (done < count) {
    u64 leaf = first_leaf + done, d = leaf / per, take = std::min<u64>(count - done, (d + 1) * per - leaf);
    P2B_TRY(p2b_batch_get_leaves(b->shard[d], leaf, take, out + done * b->info.leaf_len));
    done += take;
  }
  return P2B_OK;
}

// ---- commit from coefficients every device already holds (quotient chunks) -----------------------------------------------
// d_coeffs[d]: device d's copy of the full coefficient matrix [P][n].  Every device extends and hashes its own coset blocks;
// only the top-layer nodes are exchanged.  (PolynomialBatch::from_coeffs, fri/oracle.rs:911-977)
extern "C" int p2b_mgpu_commit_from_device_coeffs(p2b_mgpu* g, const uint64_t* const* d_coeffs, uint32_t degree_log, uint64_t num_polys,
                                                  uint32_t rate_bits, uint32_t cap_height, p2b_mgpu_batch** out) {
  if (!g || !d_coeffs || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  *out = nullptr;
  const int G = g->n;
  const u64 R = (u64)1 << rate_bits;
  if ((u64)G > R) return fail(P2B_ERR_INVALID, "%d devices exceed the 2^rate_bits = %llu coset blocks", G, (unsigned long long)R);
  p2b_mgpu_batch* mb = new (std::nothrow) p2b_mgpu_batch();
  if (!mb) return fail(P2B_ERR_OOM, "host allocation failed");
  mb->g = g;
  mb->shard.assign(G, nullptr);
  auto body = [&]() -> int {
    u32 top = 0;
    u64 count = 0;
    for (int d = 0; d < G; d++) {
      if (!d_coeffs[d]) return fail(P2B_ERR_INVALID, "NULL coefficient pointer for device %d", d);
      P2B_TRY(p2b_commit_blocks(g->ctx[d], d_coeffs[d], degree_log, num_polys, rate_bits, cap_height, nullptr, (u64)d * (R / G), R / G, &mb->shard[d]));
      top = mb->shard[d]->top_layer;
      count = mb->shard[d]->local_leaves >> top;
    }
    if (G > 1) P2B_TRY(mgpu_exchange_nodes(g, mb, top, count));
    return P2B_OK;
  };
  int rc = body();
  if (rc != P2B_OK) {
    p2b_mgpu_batch_destroy(mb);
    return rc;
  }
  mb->info = mb->shard[0]->info;
  *out = mb;
  return P2B_OK;
}

// ---- quotient polynomials over sharded batches ------------------------------------------------------------------------------
// compute_quotient_polys (plonk/prover.rs:790-1034) on the devices that hold the rows.  Point i of the quotient domain reads
// leaf row reverse_bits(i << (rate_bits - qdb)) -- held by the device whose index is reverse_bits(i mod G) -- and, for
// Z(g x), the row of point i + 2^qdb, which has the same residue mod G as long as G divides 2^qdb: every device evaluates
// exactly the points whose rows it owns, no row crosses NVLink.  The compact value vectors (2 * 8n / G words per device) are
// pushed to every peer, interleaved into the natural order and inverse-transformed on every device, so that each ends up
// with the quotient coefficients [num_challenges][lde_size] it needs to commit the quotient chunks
// (p2b_mgpu_commit_from_device_coeffs).  d_coeffs_out[d]: device d's output buffer (device pointers, caller-allocated).
extern "C" int p2b_mgpu_quotient_polys(p2b_mgpu* g, const p2b_circuit* circuit, const p2b_mgpu_batch* wires, const p2b_mgpu_batch* zs_pp,
                                       const p2b_mgpu_batch* consts_sigmas, const uint64_t* public_inputs_hash, const uint64_t* betas,
                                       const uint64_t* gammas, const uint64_t* alphas, uint64_t* const* d_coeffs_out) {
  if (!g || !circuit || !wires || !zs_pp || !consts_sigmas || !d_coeffs_out) return fail(P2B_ERR_INVALID, "NULL argument");
  const int G = g->n;
  u32 g_log = 0;
  while ((1 << g_log) < G) g_log++;
  const u32 qdf = circuit->quotient_degree_factor;
  u32 qdb = 0;
  while ((1u << qdb) < qdf) qdb++;
  if (g_log > qdb) return fail(P2B_ERR_UNSUPPORTED, "%d devices exceed the 2^quotient_degree_bits = %u point classes", G, 1u << qdb);
  const u32 nc = circuit->num_challenges;
  const u64 lde_size = (u64)1 << (circuit->degree_bits + qdb), count = lde_size >> g_log;
  for (const p2b_mgpu_batch* b : {wires, zs_pp, consts_sigmas})
    if ((int)b->shard.size() != G || b->info.degree_log != circuit->degree_bits || b->info.rate_bits != circuit->rate_bits)
      return fail(P2B_ERR_INVALID, "batch does not match the device group / circuit shape");
  std::vector<u64*> parts(G, nullptr), vals(G, nullptr);
  auto body = [&]() -> int {
    for (int d = 0; d < G; d++) {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      CUDA_TRY(pool_alloc(&parts[d], (u64)G * nc * count * sizeof(u64), c->stream));
      CUDA_TRY(pool_alloc(&vals[d], (u64)nc * lde_size * sizeof(u64), c->stream));
      CUDA_TRY(cudaEventRecord(g->ev_done[d], c->stream));
    }
    for (int s = 0; s < G; s++)
      for (int d = 0; d < G; d++) {
        CUDA_TRY(cudaSetDevice(g->ctx[s]->device));
        CUDA_TRY(cudaStreamWaitEvent(g->xfer[s], g->ev_done[d], 0));
      }
    for (int d = 0; d < G; d++) {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      const p2b_batch *w = wires->shard[d], *z = zs_pp->shard[d], *k = consts_sigmas->shard[d];
      u64 rev = 0;   // points i = reverse_bits(d) (mod G)
      for (u32 t = 0; t < g_log; t++) rev |= (((u64)d >> t) & 1) << (g_log - 1 - t);
      u64* mine = parts[d] + (u64)d * nc * count;
      P2B_TRY(quotient_impl(c, circuit, w->leaves - w->first_leaf * w->info.leaf_len, w->info.leaf_len,
                            z->leaves - z->first_leaf * z->info.leaf_len, z->info.leaf_len,
                            k->leaves - k->first_leaf * k->info.leaf_len, k->info.leaf_len, public_inputs_hash, betas, gammas, alphas, mine,
                            nullptr, nullptr, rev, (u64)G, w->first_leaf + w->local_leaves - 1));
      P2B_TRY(mgpu_events(g, d, 1));
      CUDA_TRY(cudaEventRecord(g->ev_ifft[d][0], c->stream));
      CUDA_TRY(cudaStreamWaitEvent(g->xfer[d], g->ev_ifft[d][0], 0));
      for (int t = 0; t < G; t++)
        if (t != d)
          CUDA_TRY(cudaMemcpyPeerAsync(parts[t] + (u64)d * nc * count, g->ctx[t]->device, mine, c->device, (u64)nc * count * sizeof(u64), g->xfer[d]));
      CUDA_TRY(cudaEventRecord(g->ev_nodes[d], g->xfer[d]));
    }
    for (int d = 0; d < G; d++) {
      p2b_ctx* c = g->ctx[d];
      CUDA_TRY(cudaSetDevice(c->device));
      for (int s = 0; s < G; s++)
        if (s != d) CUDA_TRY(cudaStreamWaitEvent(c->stream, g->ev_nodes[s], 0));
      quotient::interleave_parts_kernel<<<(unsigned)((lde_size + 255) / 256), 256, 0, c->stream>>>(parts[d], vals[d], lde_size, nc, g_log);
      c->launches++;
      CUDA_TRY(cudaGetLastError());
      if (!d_coeffs_out[d]) return fail(P2B_ERR_INVALID, "NULL output pointer for device %d", d);
      P2B_TRY(quotient_values_to_coeffs(c, vals[d], d_coeffs_out[d], circuit->degree_bits + qdb, nc));
    }
    return P2B_OK;
  };
  int rc = body();
  for (int d = 0; d < G; d++) {
    cudaSetDevice(g->ctx[d]->device);
    // the peers' pushes into parts[d] must have landed before it is returned to the pool: the interleave kernel on this
    // device's stream waited for them, and the frees below are ordered behind it
    if (parts[d]) cudaFreeAsync(parts[d], g->ctx[d]->stream);
    if (vals[d]) cudaFreeAsync(vals[d], g->ctx[d]->stream);
  }
  return rc;
}

// ---- FRI opening proof over sharded oracles -----------------------------------------------------------------------------------
// PolynomialBatch::prove_openings (fri/oracle.rs:1046-1110).  Everything that reads coefficients (the batch reduction, the
// commit phase, grinding) runs on the first device, whose shards hold complete coefficient copies; the FriInitialTreeProof
// rows and Merkle paths of every query come from the devices that own the leaves (p2b_mgpu_batch_open_rows).
extern "C" int p2b_mgpu_fri_prove_openings(p2b_mgpu* g, const p2b_mgpu_batch* const* oracles, uint32_t num_oracles,
                                           const p2b_fri_batch_info* batches, uint32_t num_batches, p2b_challenger* challenger,
                                           const p2b_fri_params* params, p2b_fri_proof** out) {
  if (!g || !oracles || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  std::vector<const p2b_batch*> first(num_oracles);
  for (u32 o = 0; o < num_oracles; o++) {
    if (!oracles[o] || oracles[o]->shard.empty()) return fail(P2B_ERR_INVALID, "oracle %u is NULL", o);
    first[o] = oracles[o]->shard[0];
  }
  P2B_TRY(p2b_mgpu_synchronize(g));   // the commits that produced the shards (other devices' streams) are complete
  fri_opener opener = [&](u32 o, const u64* idx, u64 Q, u64* rows, u64* sibs) -> int {
    return p2b_mgpu_batch_open_rows(oracles[o], idx, Q, rows, sibs);
  };
  return fri_prove_impl(g->ctx[0], first.data(), num_oracles, batches, num_batches, challenger, params, &opener, out);
}
// OpeningSet::new's eval_commitment (plonk/proof.rs:313-319) for one sharded batch.  Every device holds all coefficient
// columns (they were exchanged for the LDE), so the polynomials are dealt out: device d evaluates [P d / G, P (d+1) / G) from
// its own copy and writes its slice of `out`; all devices are started before the first result is waited for.
extern "C" int p2b_mgpu_eval_openings(p2b_mgpu* g, const p2b_mgpu_batch* b, const uint64_t point[2], uint64_t* out) {
  if (!g || !b || b->shard.empty() || !point || !out) return fail(P2B_ERR_INVALID, "NULL argument");
  const int G = g->n;
  const u64 P = b->shard[0]->info.num_polys;
  std::vector<u64*> d_out(G, nullptr);
  int rc = P2B_OK;
  for (int d = 0; d < G && rc == P2B_OK; d++) {
    if (!b->shard[d] || !b->shard[d]->coeffs) rc = fail(P2B_ERR_INVALID, "shard %d holds no coefficients", d);
    else rc = eval_openings_enqueue(g->ctx[d], b->shard[d], point, P * d / G, P * (d + 1) / G - P * d / G, &d_out[d]);
  }
  for (int d = 0; d < G; d++) {   // collect (and release) whatever was started, also after an error
    int r2 = eval_openings_collect(g->ctx[d], d_out[d], P * (d + 1) / G - P * d / G, out + 2 * (P * d / G));
    if (rc == P2B_OK) rc = r2;
  }
  return rc;
}
