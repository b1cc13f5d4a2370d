#!/usr/bin/env python3
"""bench.py -- LDE + Merkle commit (PolynomialBatch::from_values) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU algorithm (oracle port) on the host cores

Workload (BASELINE.json configs[1]): 2^20 rows x 135 Goldilocks columns, rate_bits = 3, cap_height = 4, Poseidon
Merkle tree, no blinding; synthetic values from splitmix64.  A step is one commit of that matrix:
batched inverse NTT -> coset LDE in leaf order -> leaf hashing -> digest layers -> cap.  At N > 1 the SAME commit is
sharded over the ranks (strong scaling): 8-column blocks dealt round-robin for the iNTT, one NCCL all-gather of the
coefficients per round overlapped with the LDE + progressive leaf hashing of the previous round, coset blocks / cap
sub-trees per rank for LDE + hashing, an all-gather of the cap entries (plonky2-gpu_b200/sharded.py).

Prints ONE JSON line (rank 0).  `value` = ms per commit with inputs resident in HBM (device events, max over ranks);
`e2e` = the same through the public API from pinned HOST buffers (H2D values, D2H coefficients + cap inside the timed
region, pipelined behind the compute in 16-column groups); `roofline` = the dominant kernel (leaf hashing) against measured HBM bandwidth, with `roofline_int` giving the
figure that actually binds it (integer issue rate); `cpu_baseline` = the oracle port on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x504C4F4E4B5932
RATE_BITS, CAP_HEIGHT = 3, 4
METRIC = "LDE+Merkle commit ms (2^20x135 cols, rate 3)"
# dynamic thread-instructions of one Poseidon permutation in the shipped SASS: ncu "Instructions Executed" per SASS line of
# hash_leaves_kernel (tools/ncu_opmix.py) divided by the permutations, profiles/r02_hash_leaves_ncu.md
INSTR_PER_PERM = 16100
WIDE_PER_PERM = 2878    # IMAD.WIDE.U32 (all forms)
FP64_PER_PERM = 5463    # DADD + DFMA
# measured issue costs per warp-instruction (tools/int_peak.cu "clean mixes", profiles/r02_pipe_model.md): IMAD.WIDE 4.36
# cycles, FP64 2.2 cycles, and the two do NOT overlap (W + kD costs 4.36 + 2.2 k), while ALU / 32-bit IMAD / XU work does
WIDE_CYCLES, FP64_CYCLES = 4.36, 2.2
MIXED_INT_LANES_PER_CLK_PER_SM = 119.6  # IADD3 + IMAD.X on both integer pipes (profiles/int_peak_r01.md)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE leaf-hash launch (2^20 leaves x 135) from the committed `ncu --set full`
# capture of this bench command (profiles/r02c_final.md); a constant from that capture, not measured in this run --
# the line says so in roofline.traffic_source
HASH_LAUNCH_DRAM_BYTES_2P20_X135 = 1258196000 + 59403520
INT_LANES_PER_CLK_PER_SM = 64  # measured IADD3 / IMAD issue rate on B200 (tools/int_peak.cu, profiles/int_peak_r01.md)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="commit", choices=["commit", "prove-ecc", "prove-recursion"],
                    help="commit = the headline (BASELINE.json configs[1], or configs[4] sizes with --n-log/--polys); prove-* = the "
                         "prove() data path on the configs[3] / configs[2] shapes (single GPU)")
    ap.add_argument("--n-log", type=int, default=20, help="log2 rows (development override; the judged run uses 20)")
    ap.add_argument("--polys", type=int, default=135)
    ap.add_argument("--cpu-sample-log", type=int, default=0,
                    help="log2 rows of the bounded CPU sample (0 = largest power of two whose run fits --cpu-budget-s)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-clock budget of all CPU-arm steps together")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA kernels (oracle/_ref)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self._stop_evt, self.proc = gpu_index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [x for x in sm if x > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's CPU path, on all host threads
# ---------------------------------------------------------------------------------------------------------------
def cpu_threads():
    """All host threads, set explicitly: torchrun exports OMP_NUM_THREADS=1, which would silently make this a 1-core run."""
    import oracle
    oracle.build()
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    oracle.set_threads(n)
    return oracle.get_threads()


def pick_cpu_sample_log(args, total_reps):
    """Largest sample (<= the workload itself) whose `total_reps` repetitions fit the CPU budget, from a 2^13-row probe."""
    if args.cpu_sample_log:
        return min(args.cpu_sample_log, args.n_log)
    probe = min(13, args.n_log)
    ms, _ = cpu_commit_ms(probe, args.polys)
    ms, _ = cpu_commit_ms(probe, args.polys)
    k = probe
    while k < args.n_log and (ms * (1 << (k + 1 - probe)) * 1.1e-3) * total_reps <= args.cpu_budget_s:
        k += 1
    return k


def cpu_commit_ms(sample_log, polys, reps=1):
    """Times PolynomialBatch::from_values of the oracle on a 2^sample_log x polys sample of the workload."""
    import oracle
    cpu_threads()
    rng = np.random.default_rng(SEED & 0xFFFFFFFF)
    values = rng.integers(0, oracle.ORDER, size=(polys, 1 << sample_log), dtype=np.uint64)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        oracle.batch_from_values(values, RATE_BITS, CAP_HEIGHT, want_leaves=True, want_digests=True)
        dt = (time.perf_counter() - t) * 1e3
        best = dt if best is None else min(best, dt)
    return best, oracle.get_threads()


def workload_name(n_log, polys):
    return ("PolynomialBatch::from_values 2^%d x %d Goldilocks, rate_bits 3, cap_height 4, Poseidon Merkle (%s)"
            % (n_log, polys, "BASELINE.json configs[1]" if (n_log, polys) == (20, 135)
               else "BASELINE.json configs[4] scale sweep: NOT the headline size the metric name quotes"))


def cpu_baseline_obj(args, value_ms, cores, sample_log, kind="port"):
    scale = 1 << (args.n_log - sample_log)
    if scale == 1:
        how = "the whole workload, measured (no scaling)"
    else:
        how = ("commit of 2^%d x %d measured, then scaled x%d to 2^%d rows (work is linear in rows up to the log factor of the "
               "NTT, which is <10%% of the time)" % (sample_log, args.polys, scale, args.n_log))
    return {"value": value_ms, "unit": "ms", "cores": cores, "kind": kind, "sample_rows_log2": sample_log,
            "sample": "oracle (C/OpenMP restatement of the reference CPU path, all %d host threads): %s" % (cores, how)}


def run_reference(args, rank):
    if rank != 0:
        return
    sample_log = pick_cpu_sample_log(args, args.steps + args.warmup)
    scale = 1 << (args.n_log - sample_log)
    for _ in range(args.warmup):
        cpu_commit_ms(sample_log, args.polys)
    times, cores = [], 1
    for _ in range(args.steps):
        ms, cores = cpu_commit_ms(sample_log, args.polys)
        times.append(ms * scale)
    ms = sum(times) / len(times)
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args.n_log, args.polys),
                       "timing": "host wall clock of the CPU run (2^%d-row sample x%d)" % (sample_log, scale)},
            "cpu_baseline": cpu_baseline_obj(args, ms, cores, sample_log),
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def headline_golden_cap(n_log, polys):
    if (n_log, polys) != (20, 135):
        return None
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))
        return np.array(g["headline"]["cap"], dtype=np.uint64)
    except Exception:
        return None


def parity_preflight(p2b, sharded, ctx, engine, comm, commit_sharded, rank, world, P, pipelined, n_log=12):
    """Before any timing: commit a 2^12 x P slice of the synthetic matrix through the code path about to be timed and compare
    with the CPU oracle -- at N = 1 coefficients, leaves, digests and cap; at N > 1 this rank's leaf rows, its opened rows
    + Merkle paths (verified against the oracle's cap) and the cap.  Raises on any mismatch."""
    import torch
    import oracle
    oracle.build()
    n = 1 << n_log
    N = n << RATE_BITS
    L = p2b.lib()
    full = p2b.DeviceBuffer(ctx, P * n)
    ctx.fill_synthetic(full, P * n, SEED)
    ctx.synchronize()
    values = full.to_host(P * n).reshape(P, n)
    want = oracle.batch_from_values(values, RATE_BITS, CAP_HEIGHT)
    if world == 1:
        b = p2b.PolynomialBatch.from_values(ctx, values, RATE_BITS, CAP_HEIGHT)
        ok = (np.array_equal(b.cap(), want.cap) and np.array_equal(b.polynomials(), want.coeffs)
              and np.array_equal(b.leaves(), want.leaves) and np.array_equal(b.digests(), want.digests))
        b.close()
        full.free()
        if not ok:
            raise SystemExit("bench.py: parity preflight failed (2^%d x %d commit differs from the CPU oracle)" % (n_log, P))
        return "2^%d x %d: coefficients, leaves, digests, cap == CPU oracle" % (n_log, P)
    if pipelined:
        shard = sharded.pack_local(values, P, world, rank)
    else:
        c0, c1, cmax = sharded.column_shard(P, world, rank)
        shard = np.zeros((cmax, n), dtype=np.uint64)
        shard[: c1 - c0] = values[c0:c1]
    t = torch.from_numpy(shard.view(np.int64)).cuda()
    b = commit_sharded(engine, comm, t, P, n_log, RATE_BITS, CAP_HEIGHT)
    engine.synchronize()
    per = N // world
    ok = np.array_equal(b.cap(), want.cap)
    mine = b.leaves()
    ok = ok and np.array_equal(np.asarray(mine).reshape(-1, P)[:per], want.leaves[rank * per:(rank + 1) * per])
    idx = [int(x) for x in np.random.default_rng(5).integers(0, N, size=16)]
    rows, sibs = sharded.sharded_open_rows(engine, comm, b, idx, n_log, RATE_BITS, CAP_HEIGHT, P)
    for k, x in enumerate(idx):
        ok = ok and np.array_equal(rows[k], want.leaves[x]) and oracle.merkle_verify(rows[k], x, want.cap, sibs[k])
    b.close()
    full.free()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    torch.distributed.all_reduce(flag)
    if int(flag[0]) != 0:
        raise SystemExit("bench.py: rank %d: parity preflight failed (sharded 2^%d x %d commit differs from the CPU oracle)" % (rank, n_log, P))
    return "2^%d x %d sharded over %d ranks: each rank's leaf rows, 16 opened rows + Merkle paths, cap == CPU oracle" % (n_log, P, world)


def run_b200(args, rank, world, local_rank):
    import torch
    import plonky2_gpu_b200 as p2b
    from plonky2_gpu_b200 import sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p2b.build()
    ctx = p2b.Context(local_rank)
    L = p2b.lib()
    n_log, P = args.n_log, args.polys
    n, N = 1 << n_log, 1 << (n_log + RATE_BITS)
    # ---- inputs resident in HBM: this rank's columns of the synthetic value matrix ----
    # N = 1: all P columns.  N > 1: the rank's slice of every exchange round of the pipelined flow (sharded.local_layout:
    # round j = consecutive columns [col0_j, col0_j + width_j), of which rank r holds `per_j` consecutive ones at rows
    # [row0_j, row0_j + per_j) of its local buffer; zero rows where a round is ragged).
    pipelined = world > 1 and P > 4 and not os.environ.get("P2B_BENCH_UNPIPELINED")  # env knob: A/B against the one-shot exchange
    if pipelined:
        cmax, layout = sharded.local_layout(P, world, rank)
        col_ranges = [(row0, a_, b_) for row0, a_, b_ in layout if b_ > a_]   # (local row, c0, c1)
    else:
        c0, c1, cmax = sharded.column_shard(P, world, rank)
        col_ranges = [(0, c0, c1)] if c1 > c0 else []
    own_cols = sum(b_ - a_ for _, a_, b_ in col_ranges)
    vals = torch.zeros((cmax, n), dtype=torch.int64, device="cuda")
    work = torch.empty_like(vals)
    torch.cuda.synchronize()
    for row, a_, b_ in col_ranges:
        p2b._check(L.p2b_fill_synthetic(ctx.handle, vals.data_ptr() + row * n * 8, (b_ - a_) * n, SEED, a_ * n))
    ctx.synchronize()
    engine = sharded.GpuEngine(ctx)
    comm = sharded.TorchComm(dist) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    commit_sharded = sharded.sharded_commit_from_values_pipelined if pipelined else sharded.sharded_commit_from_values

    def step_device():
        if world == 1:
            b = p2b.PolynomialBatch.from_values(ctx, (_Ptr(vals.data_ptr()), P, n), RATE_BITS, CAP_HEIGHT)
        else:
            work.copy_(vals)  # the sharded path transforms its columns in place
            b = commit_sharded(engine, comm, work, P, n_log, RATE_BITS, CAP_HEIGHT)
        engine.synchronize()
        return b

    class _Ptr:  # minimal DeviceBuffer stand-in for torch-owned memory
        def __init__(self, ptr):
            self.ptr = ptr

    # ---- parity preflight (never timed): the same code path on a 2^12-row slice against the CPU oracle ----
    preflight = parity_preflight(p2b, sharded, ctx, engine, comm, commit_sharded, rank, world, P, pipelined)

    # ---- warm-up ----
    cap_check = None
    for _ in range(max(args.warmup, 3)):
        b = step_device()
        cap_check = b.cap()
        b.close()
    barrier()
    # the full cap of the headline matrix as computed by the CPU oracle (tests/golden/oracle_golden.json, generated by
    # tests/golden/make_golden.py --headline); every rank must reproduce all 16 x 4 words or the run is void
    golden_cap = headline_golden_cap(n_log, P)
    if golden_cap is not None:
        if not np.array_equal(np.asarray(cap_check, dtype=np.uint64), golden_cap):
            raise SystemExit("bench.py: rank %d: cap of the headline commit differs from the CPU-oracle golden -- run void" % rank)
        cap_status = "all 16x4 words equal the CPU-oracle golden on every rank"
    else:
        cap_status = "no golden for this size (preflight only)"

    # ---- timed: device-resident ----
    gpu_index = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank]) if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank
    sampler = ClockSampler(gpu_index) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    ctx.time_leaf_hash(True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        b = step_device()
        b.close()
    dev_ms = ctx.timer_stop_ms()
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    hash_ms, hash_launches = ctx.leaf_hash_time()
    ctx.time_leaf_hash(False)
    launches = ctx.launch_count - launches0
    if sampler:
        sampler.stop()
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms / args.steps

    # ---- timed: end to end from pinned host buffers through the public API ----
    host_vals = p2b.PinnedBuffer(max(cmax, 1) * n)
    host_coef = p2b.PinnedBuffer(max(cmax, 1) * n)
    host_vals.array[:] = vals.cpu().numpy().view(np.uint64).reshape(-1)
    hv = host_vals.array.reshape(cmax, n)
    hc = host_coef.array.reshape(cmax, n)
    cap_host = np.empty((1 << CAP_HEIGHT, 4), dtype=np.uint64)
    copy_stream = torch.cuda.Stream()
    host_coef_t = torch.from_numpy(host_coef.array.view(np.int64))  # same pinned memory, as torch tensors for the async copies
    host_vals_t = torch.from_numpy(host_vals.array.view(np.int64)).view(cmax, n)

    def step_e2e():
        if world == 1:
            # H2D of the values and D2H of the coefficients (the reference keeps them host-side, oracle.rs:403-407)
            # happen inside the call, overlapped with the transforms / the tree
            b = p2b.PolynomialBatch.from_values(ctx, hv[:P], RATE_BITS, CAP_HEIGHT, coeffs_out=hc[:P])
            b.cap(out=cap_host)                                                       # D2H result (synchronises)
            b.close()
        else:
            def coeffs_to_host(shard):
                # each rank returns the coefficient columns it transformed, on a side stream while the exchange, the LDE
                # and the tree run (the single-GPU call does the same inside p2b_commit_from_values_ex)
                copy_stream.wait_stream(torch.cuda.current_stream())   # the engine ordered the transforms before this stream
                with torch.cuda.stream(copy_stream):
                    for row, a_, b_ in col_ranges:
                        host_coef_t[row * n: (row + b_ - a_) * n].copy_(shard.view(-1)[row * n: (row + b_ - a_) * n], non_blocking=True)

            b = commit_sharded(engine, comm, work, P, n_log, RATE_BITS, CAP_HEIGHT, on_coeffs_ready=coeffs_to_host, host_values=host_vals_t)
            b.cap(out=cap_host)
            copy_stream.synchronize()
            b.close()
        engine.synchronize()

    for _ in range(2):
        step_e2e()
    if world > 1:  # the overlapped copy delivered this rank's coefficient columns
        for row, a_, b_ in col_ranges:
            assert torch.equal(host_coef_t[row * n: (row + b_ - a_) * n], work.view(-1)[row * n: (row + b_ - a_) * n].cpu()), "coefficient D2H mismatch"
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    w0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_dev = ctx.timer_stop_ms()
    barrier()
    e2e_wall = (time.perf_counter() - w0) * 1e3
    t = torch.tensor([e2e_dev, e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[1]) / e2e_steps  # host wall clock: the region contains host-side copies and API calls
    h2d_bytes = P * n * 8
    d2h_bytes = P * n * 8 + (1 << CAP_HEIGHT) * 32

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda and (n_log, P) == (20, 135):
        # GPU-vs-GPU baseline: the reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100a) on the same box
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ref_cuda_bench
            # the reference kernels printf their own timings: keep them off this process's stdout (ONE JSON line)
            sys.stdout.flush()
            saved_fd = os.dup(1)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            try:
                r = ref_cuda_bench.measure(n_log, P, reps=2, ctx=ctx, with_ours=False)
            finally:
                import ctypes
                ctypes.CDLL(None).fflush(None)
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
                os.close(devnull)
            if "ref_ms" in r:
                ref_cuda = {"value": r["ref_ms"], "unit": "ms", "kind": "reference CUDA (cuda/plonky2_gpu.cu ifft + merkle_tree_from_coeffs, "
                            "recompiled unmodified for sm_100a), same B200, device-resident, CUDA events",
                            "cap_equal": r["ref_cap_word0"] == "%016x" % int(cap_check[0][0])}
            else:
                ref_cuda = r
        except Exception as e:  # the baseline must never take the run down
            ref_cuda = {"unavailable": repr(e)[:200]}
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        clocks = sampler.summary() if sampler else {}
        # roofline of the dominant kernel (leaf hashing): algorithmic bytes of one launch = the rows of one coset block
        # read once + their digests written once (8*leaf_len per leaf in, 32 B per leaf out)
        # N = 1 (and the unpipelined flow): one launch per coset block, n leaves each.  Pipelined multi-GPU flow: the leaf
        # hashing of the rank's N/world leaves is split over one absorb_columns_kernel launch per exchange round; a launch's
        # algorithmic bytes / permutations are the per-step totals divided by the launches per step.
        launches_per_step = max(hash_launches, 1) / max(args.steps, 1)
        local_leaves = N // world
        alg_bytes = local_leaves * (P * 8 + 32) / launches_per_step
        avg_hash_ms = hash_ms / max(hash_launches, 1)
        achieved_gbs = alg_bytes / (avg_hash_ms * 1e-3) / 1e9 if avg_hash_ms > 0 else 0.0
        perms_per_launch = local_leaves * (-(-P // 8)) / launches_per_step
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        int_peak = INT_LANES_PER_CLK_PER_SM * 148 * sm_mhz * 1e6          # thread-instructions / s
        int_ach = perms_per_launch * INSTR_PER_PERM / (avg_hash_ms * 1e-3) if avg_hash_ms > 0 else 0.0
        perm_rate = perms_per_launch / (avg_hash_ms * 1e-3) if avg_hash_ms > 0 else 0.0
        heavy_cycles = WIDE_PER_PERM * WIDE_CYCLES + FP64_PER_PERM * FP64_CYCLES       # per warp (32 permutations), per sub-partition
        heavy_peak = 148 * 4 * sm_mhz * 1e6 * 32 / heavy_cycles                         # permutations / s if that resource never idled
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(n_log, P),
                       "sharding": ("columns dealt in exchange rounds of 8, 8, 16, .. 8N consecutive columns for the iNTT, one NCCL all-gather per round "
                                    "overlapped with LDE + progressive leaf hashing of the previous round, coset blocks / cap sub-trees per rank") if pipelined
                       else ("columns for iNTT, coset blocks / cap sub-trees for LDE+Merkle" if world > 1 else "single GPU"),
                       "l2": "inputs_exceed_l2 (values %.2f GB, LDE %.2f GB per step vs 126 MB L2)" % (P * n * 8 / 1e9, P * N * 8 / 1e9),
                       "timing": "CUDA events on the library stream around all steps, max over ranks; wall %.1f ms/step" % (wall_ms / args.steps),
                       "cap_word0": "%016x" % int(cap_check[0][0]), "cap_check": cap_status, "parity_preflight": preflight},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "note": "pinned host values -> p2b_commit_from_values -> D2H coefficients + cap; host wall clock, max over ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "merkle::absorb_columns_kernel (leaf hashing, one launch per exchange round)" if pipelined else "merkle::hash_leaves_kernel", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                         "traffic": HASH_LAUNCH_DRAM_BYTES_2P20_X135 if (n_log == 20 and P == 135 and not pipelined) else None, "traffic_source": "ncu --set full capture of this command, profiles/r02c_final.md (constant, not re-measured per run)",
                         "peak_source": peak_kind + " (burst copy bandwidth)",
                         "launches_timed": hash_launches, "avg_launch_ms": avg_hash_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "share_of_step": hash_ms / max(dev_ms, 1e-9)},
            "roofline_int": {"bound": "IMAD.WIDE + FP64 issue cycles (mutually exclusive on sm_100a, measured): the binding resource of "
                                      "the Poseidon kernel", "achieved": perm_rate / 1e6, "peak": heavy_peak / 1e6, "unit": "Mperm/s",
                             "frac": perm_rate / heavy_peak if heavy_peak else None,
                             "cycles_per_warp_permutation": heavy_cycles, "wide_per_permutation": WIDE_PER_PERM,
                             "fp64_per_permutation": FP64_PER_PERM, "instr_per_permutation": INSTR_PER_PERM, "sm_mhz": sm_mhz,
                             "thread_instr_per_s_T": int_ach / 1e12,
                             "frac_of_64_lane_single_pipe": int_ach / int_peak if int_peak else None,
                             "frac_of_mixed_alu_fma_119.6_lanes": int_ach / (int_peak * MIXED_INT_LANES_PER_CLK_PER_SM / INT_LANES_PER_CLK_PER_SM) if int_peak else None},
            "clocks": clocks,
        }
        if ref_cuda is not None:
            line["ref_cuda_baseline"] = ref_cuda
        if not args.no_cpu_baseline:
            args.cpu_budget_s = min(args.cpu_budget_s, 40.0)
            sample_log = pick_cpu_sample_log(args, 2)
            ms, cores = cpu_commit_ms(sample_log, P, reps=2)
            line["cpu_baseline"] = cpu_baseline_obj(args, ms * (1 << (n_log - sample_log)), cores, sample_log)
        print(json.dumps(line), flush=True)
    engine.synchronize()
    torch.cuda.synchronize()
    host_vals.free()
    host_coef.free()
    del engine
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_prove(args):
    """--workload prove-ecc / prove-recursion: BASELINE.json's "e2e prove s" on the config 4 / config 3 shapes -- every stage of
    prove() between the witness and the proof's field elements, on the device (plonky2-gpu_b200/pipeline.py).  One JSON line."""
    import torch
    import plonky2_gpu_b200 as p2b
    from plonky2_gpu_b200.pipeline import ProvePipeline, STAGES
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    kind = "ecc" if args.workload == "prove-ecc" else "recursion"
    n_log = args.n_log if any(a.startswith("--n-log") for a in sys.argv) else (17 if kind == "ecc" else 16)
    p2b.build()
    ctx = p2b.Context(0)
    # parity preflight on a 2^8-row instance of the same gate set: quotient values at sampled points == CPU oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.test_gpu_configs import _oracle_circuit
    from oracle import quotient as Q
    import oracle
    small = ProvePipeline(ctx, kind, 8)
    _, keep = small.prove(keep=True)
    circ = _oracle_circuit(small)
    pts = [0, 1, 77, small.size - 1]
    need = sorted(set(r for i in pts for r in Q.quotient_point_rows(circ, i)))
    rows = {name: dict(zip(need, b.open_rows(need, with_proofs=False)[0])) for name, b in (("w", keep["b_w"]), ("z", keep["b_z"]), ("cs", small.b_cs))}
    want = Q.compute_quotient_values(circ, rows["w"], rows["z"], rows["cs"], small.pih, small.betas, small.gammas, small.alphas, points=pts)
    qv = keep["quotient_values"].to_host(small.nc * small.size).reshape(small.nc, small.size)
    if any(int(qv[c][i]) != wv[c] for i, wv in zip(pts, want) for c in range(small.nc)):
        raise SystemExit("bench.py: parity preflight failed (quotient values differ from the CPU oracle)")
    keep["proof"].close()
    for k in ("b_w", "b_z", "b_q"):
        keep[k].close()
    small.close()
    pipe = ProvePipeline(ctx, kind, n_log)
    for _ in range(max(args.warmup, 3)):
        pipe.prove()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    times, walls = {}, []
    launches0 = ctx.launch_count
    for _ in range(args.steps):
        walls.append(pipe.prove(times))
    launches = ctx.launch_count - launches0
    sampler.stop()
    ms = sum(walls) / len(walls)
    line = {"metric": "e2e prove data path ms (%s)" % pipe.describe(), "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "prove() data path, BASELINE.json configs[%d]: %s" % (3 if kind == "ecc" else 2, pipe.describe()),
                       "timing": "host wall clock per proof (the stages synchronise between each other); stage figures are CUDA events",
                       "parity_preflight": "2^8-row instance of the same gate set: quotient values at sampled points == CPU oracle",
                       "out_of_scope": "circuit building, witness generation, the per-circuit constants_sigmas commit"},
            "stages_ms": {k: sum(times[k]) / len(times[k]) for k in STAGES if k in times},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "the witness is generated on the device in this flow (no host copy on the path); proof bytes are read back inside the FRI call"},
            "gpu_launches": launches, "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)
    pipe.close()
    ctx.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not (world == 1 and args.impl == "reference"):
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                             "--master-addr 127.0.0.1 --master-port 29501 bench.py --gpus %d ..." % (args.gpus, args.gpus, args.gpus))
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload != "commit":
        if rank == 0:
            run_prove(args)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
