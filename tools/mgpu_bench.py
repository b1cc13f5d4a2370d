"""Single-process multi-device commit (p2b_mgpu_*) timed on G devices of one node: the headline commit, device-resident
(CUDA events on every device, max) and end to end from pinned host memory (host wall clock).
    python tools/mgpu_bench.py [G=8] [n_log=20] [P=135] [steps=5]
Prints one JSON line; the cap is compared with tests/golden/oracle_golden.json for the headline shape."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import plonky2_gpu_b200 as p2b  # noqa: E402

SEED = 0x504C4F4E4B5932
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_log = int(sys.argv[2]) if len(sys.argv) > 2 else 20
P = int(sys.argv[3]) if len(sys.argv) > 3 else 135
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
n = 1 << n_log
p2b.build()
L = p2b.lib()
mg = p2b.MultiGpu(count=G)


from plonky2_gpu_b200 import sharded  # noqa: E402  (same schedule as mgpu_schedule in csrc/mgpu.cuh)


def fill_resident():
    for d in range(G):
        ptr, rounds = mg.resident_cols(d, n_log, P)
        ctxh = L.p2b_mgpu_ctx(mg.handle, d)
        rows, layout = sharded.local_layout(P, G, d)
        assert len(layout) == rounds
        for row0, c0, c1 in layout:
            if c1 > c0:
                p2b._check(L.p2b_fill_synthetic(ctxh, ptr + row0 * n * 8, (c1 - c0) * n, SEED, c0 * n))


def step_resident():
    fill_resident()
    mg.timer_start()
    b = mg.commit_resident(n_log, P, 3, 4)
    ms = mg.timer_stop_ms()
    return b, ms


golden = None
if (n_log, P) == (20, 135):
    golden = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))["headline"]["cap"], dtype=np.uint64)
cap = None
for _ in range(3):
    b, _ms = step_resident()
    cap = b.cap()
    b.close()
if golden is not None:
    assert np.array_equal(cap, golden), "cap differs from the CPU-oracle golden"
ts = []
for _ in range(steps):
    b, ms = step_resident()
    ts.append(ms)
    b.close()
mg.synchronize()
# end to end from pinned host values (+ coefficients back to the host)
hv, hc = p2b.PinnedBuffer(P * n), p2b.PinnedBuffer(P * n)
tmp = p2b.DeviceBuffer(p2b.Context(0), 8 * n) if False else None
ctx0 = L.p2b_mgpu_ctx(mg.handle, 0)
import ctypes as C  # noqa: E402
dptr = C.c_void_p()
p2b._check(L.p2b_malloc(ctx0, P * n * 8, C.byref(dptr)))
p2b._check(L.p2b_fill_synthetic(ctx0, dptr, P * n, SEED, 0))
p2b._check(L.p2b_memcpy_d2h(ctx0, hv.array.ctypes.data, dptr, P * n * 8))
p2b._check(L.p2b_free(ctx0, dptr))
vals = hv.array.reshape(P, n)
coef = hc.array.reshape(P, n)
for _ in range(2):
    b = mg.commit_from_values(vals, 3, 4, coeffs_out=coef)
    cap2 = b.cap()
    b.close()
assert np.array_equal(cap2, cap)
mg.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    b = mg.commit_from_values(vals, 3, 4, coeffs_out=coef)
    cap2 = b.cap()     # waits for the whole commit incl. the coefficient copies
    b.close()
e2e = (time.perf_counter() - t0) * 1e3 / steps
print(json.dumps({"tool": "mgpu_bench (single process, p2b_mgpu_*)", "n_gpus": G, "workload": "2^%d x %d, rate 3, cap 4" % (n_log, P),
                  "device_ms": sum(ts) / len(ts), "device_ms_min": min(ts), "e2e_ms": e2e, "peer_access": mg.peer_access,
                  "cap_word0": "%016x" % int(cap[0][0]), "cap_equals_golden": golden is not None}))
hv.free()
hc.free()
mg.close()
