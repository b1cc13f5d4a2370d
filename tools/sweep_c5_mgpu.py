"""BASELINE config 5 -- scale sweep: commit of 2^22..2^24 rows x 135..400 columns sharded over G devices of one node, driven
by ONE process through the C ABI (p2b_mgpu_*).  Device-resident input (each device's column shard filled in place), CUDA
events on every device (max), 3 warm-ups + `steps` timed commits per point.
    python tools/sweep_c5_mgpu.py [G=8] [steps=3] [rows_logs=22,23,24] [cols=135,234,400]
Per point one JSON line: ms, SHA-256 of the cap (equal across G = results independent of the sharding), and the checks run on
it: Merkle paths of opened rows verified with the CPU oracle at every point; at 2^22 rows two columns of the opened rows are
compared with the oracle's evaluation of the CPU-inverse-transformed column (LDE value parity at sweep size).
Points whose per-device footprint exceeds the 180 GB of a B200 are reported as skipped."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import plonky2_gpu_b200 as p2b  # noqa: E402
import oracle  # noqa: E402
from plonky2_gpu_b200 import sharded  # noqa: E402

SEED = 0x504C4F4E4B5932
P_FIELD = oracle.ORDER if hasattr(oracle, "ORDER") else 0xFFFFFFFF00000001
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows_logs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "22,23,24").split(",")]
cols = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "135,234,400").split(",")]
RATE, CAP = 3, 4
HBM_BUDGET = 168e9

p2b.build()
oracle.build()
L = p2b.lib()
mg = p2b.MultiGpu(count=G)


def footprint(n_log, P):
    n, N = 1 << n_log, 1 << (n_log + RATE)
    per_dev = P * n * 8 * (1 + 1.0 / G)        # all coefficients + this device's resident value columns
    per_dev += N * P * 8 / G                   # its leaf rows
    per_dev += 2 * N * 32 + 12 * 8 * N / G     # digest buffer in the whole tree's layout + sponge states
    per_dev += 2 * 64 * n * 8 * 2              # scratch of the widest exchange round (64 columns), with slack
    return per_dev


def fill(n_log, P):
    n = 1 << n_log
    for d in range(G):
        ptr, rounds = mg.resident_cols(d, n_log, P)
        ctxh = L.p2b_mgpu_ctx(mg.handle, d)
        for row0, c0, c1 in sharded.local_layout(P, G, d)[1]:
            if c1 > c0:
                p2b._check(L.p2b_fill_synthetic(ctxh, ptr + row0 * n * 8, (c1 - c0) * n, SEED, c0 * n))


def column_on_host(n_log, c):
    n = 1 << n_log
    ctx0 = mg.context(0)
    buf = p2b.DeviceBuffer(ctx0, n)
    ctx0.fill_synthetic(buf, n, SEED, c * n)
    out = np.empty(n, dtype=np.uint64)
    p2b._check(L.p2b_memcpy_d2h(ctx0.handle, out.ctypes.data, buf.ptr, n * 8))
    buf.free()
    return out


# largest first: the stream-ordered pools then serve every later (smaller) point from memory they already hold
for n_log, P in sorted(((a, b) for a in rows_logs for b in cols), key=lambda t: -(t[1] << t[0])):
    if True:
        tag = {"tool": "sweep_c5_mgpu (one process, p2b_mgpu_*)", "n_gpus": G, "rows_log2": n_log, "columns": P, "rate_bits": RATE, "cap_height": CAP,
               "lde_gb": round((1 << (n_log + RATE)) * P * 8 / 1e9, 1)}
        need = footprint(n_log, P)
        if need > HBM_BUDGET:
            print(json.dumps(dict(tag, skipped="per-device footprint %.0f GB exceeds one B200" % (need / 1e9))), flush=True)
            continue
        n, N = 1 << n_log, 1 << (n_log + RATE)
        try:
            ts = []
            b = None
            for it in range(3 + steps):
                if b is not None:
                    b.close()
                fill(n_log, P)
                mg.timer_start()
                b = mg.commit_resident(n_log, P, RATE, CAP)
                ms = mg.timer_stop_ms()
                if it >= 3:
                    ts.append(ms)
            cap = b.cap()
            rng = np.random.default_rng(n_log * 1000 + P)
            idx = sorted(set([0, N - 1, N // G - 1, N // G % N, (5 * n + 17) % N] + [int(x) for x in rng.integers(0, N, size=3)]))
            rows, sibs = b.open_rows(idx)
            paths_ok = all(oracle.merkle_verify(r, i, cap, s) for r, s, i in zip(rows, sibs, idx))
            checks = {"merkle_paths_verified_by_oracle": len(idx) if paths_ok else 0}
            if n_log <= 22:
                wN = oracle.primitive_root_of_unity(n_log + RATE)
                ok = True
                for c in (0, P - 1):
                    coeffs = oracle.ifft(column_on_host(n_log, c))
                    for Lf in (idx[1], idx[-1]):
                        x = 7 * oracle.exp(wN, oracle.reverse_bits(Lf, n_log + RATE)) % P_FIELD
                        ok &= int(rows[idx.index(Lf)][c]) == int(oracle.naive_coset_eval(coeffs, 0, int(x))[0])
                checks["lde_values_equal_oracle"] = 4 if ok else 0
            b.close()
            mg.synchronize()
            ok_all = paths_ok and checks.get("lde_values_equal_oracle", 1) != 0
            print(json.dumps(dict(tag, ms=round(sum(ts) / len(ts), 2), ms_min=round(min(ts), 2), steps=steps,
                                  cap_sha256=hashlib.sha256(cap.tobytes()).hexdigest()[:16], cap_word0="%016x" % int(cap[0][0]),
                                  checks=checks, ok=bool(ok_all))), flush=True)
            if not ok_all:
                sys.exit(1)
        except p2b.P2BError as e:
            print(json.dumps(dict(tag, error=str(e)[:200])), flush=True)
mg.close()
