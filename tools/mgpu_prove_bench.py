"""The prove() data path (plonky2-gpu_b200/pipeline.py) over G devices driven by ONE process (p2b_mgpu_*): wires commit,
Z / partial products, their commit, sharded quotient polynomials, quotient-chunk commit, openings, FRI opening proof.
    python tools/mgpu_prove_bench.py [G=8] [kind=recursion] [n_log=20] [reps=3]
Stage times are host wall clock with a group synchronisation after every stage.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import plonky2_gpu_b200 as p2b  # noqa: E402
from plonky2_gpu_b200 import sharded  # noqa: E402
from plonky2_gpu_b200.pipeline import SHAPES  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "recursion"
n_log = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sh = SHAPES[kind]
rate_bits, cap_height, nc, qdf, pow_bits, queries = 3, 4, 2, 8, 16, 28
num_wires, num_routed = sh["num_wires"], sh["num_routed"]
num_constants = len(sh["groups"]) + sh["num_gate_consts"]
n = 1 << n_log
arity, d = [], n_log
while d > 5 and d + rate_bits - 4 >= cap_height:
    arity.append(4)
    d -= 4
p2b.build()
L = p2b.lib()
mg = p2b.MultiGpu(count=G)
ctx0 = mg.context(0)
k_is = [pow(7, j, p2b.ORDER) for j in range(num_routed)]
circ = p2b.Circuit(sh["gates"], sh["sel"], sh["groups"], num_wires, num_routed, num_constants, k_is, n_log, rate_bits, nc, qdf)
rng = np.random.default_rng(0)
rnd = lambda k: [int(x) for x in rng.integers(0, p2b.ORDER, size=k, dtype=np.uint64)]   # noqa: E731
pih, betas, gammas, alphas = rnd(4), rnd(nc), rnd(nc), rnd(nc)


def fill_resident(cols, seed):
    for dev in range(G):
        ptr, rounds = mg.resident_cols(dev, n_log, cols)
        for row0, c0, c1 in sharded.local_layout(cols, G, dev)[1]:
            if c1 > c0:
                p2b._check(L.p2b_fill_synthetic(mg._ctx(dev), ptr + row0 * n * 8, (c1 - c0) * n, seed, c0 * n))


# per-circuit data (outside the timed region): constants_sigmas commit; the witness values also live on device 0 for Z
fill_resident(num_constants + num_routed, 3)
b_cs = mg.commit_resident(n_log, num_constants + num_routed, rate_bits, cap_height)
d_sigma = p2b.DeviceBuffer(ctx0, num_routed * n)
ctx0.fill_synthetic(d_sigma, num_routed * n, 3, num_constants * n)
d_wires0 = p2b.DeviceBuffer(ctx0, num_wires * n)
ctx0.fill_synthetic(d_wires0, num_wires * n, 1)
mg.synchronize()
zeta = (0x123456789abcdef, 0xfedcba987654321)
g = pow(1753635133440165772, 1 << (32 - n_log), p2b.ORDER)
zeta_next = (zeta[0] * g % p2b.ORDER, zeta[1] * g % p2b.ORDER)


def prove(times):
    def stage(name, fn):
        t = time.perf_counter()
        r = fn()
        mg.synchronize()
        times.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
        return r
    fill_resident(num_wires, 1)      # the witness columns on their owner devices (not timed: witness generation is out of scope)
    mg.synchronize()
    t_all = time.perf_counter()
    b_w = stage("commit wires", lambda: mg.commit_resident(n_log, num_wires, rate_bits, cap_height))
    zs, shape = stage("Z + partial products", lambda: p2b.partial_products_and_zs(ctx0, (d_wires0, num_wires, n), (d_sigma, num_routed, n), k_is, betas, gammas, qdf))
    b_z = stage("commit Z/pp", lambda: mg.commit_from_device_values(0, zs.ptr, n_log, shape[0], rate_bits, cap_height))
    ptrs = stage("quotient polys", lambda: mg.quotient_polys(circ, b_w, b_z, b_cs, pih, betas, gammas, alphas))
    b_q = stage("commit quotient chunks", lambda: mg.commit_from_device_coeffs(ptrs, n_log, nc * qdf, rate_bits, cap_height))
    oracles = [b_cs, b_w, b_z, b_q]

    def openings():
        for o in oracles:
            mg.eval_openings(o, zeta)
        mg.eval_openings(b_z, zeta_next)
    stage("openings", openings)
    polys = (num_constants + num_routed, num_wires, shape[0], nc * qdf)
    all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
    ch = p2b.Challenger(list(range(1, 13)), [5, 6, 7])
    pr = stage("FRI prove_openings", lambda: mg.fri_prove_openings(oracles, [(zeta, all_polys), (zeta_next, [(2, p) for p in range(nc)])], ch,
                                                                   n_log, rate_bits, cap_height, pow_bits, queries, arity))
    wall = (time.perf_counter() - t_all) * 1e3
    cap = b_w.cap()
    pr.close()
    mg.free_device_ptrs(ptrs)
    for b in (b_w, b_z, b_q):
        b.close()
    return wall, cap


prove({})
times, walls = {}, []
for _ in range(reps):
    w, cap = prove(times)
    walls.append(w)
print(json.dumps({"tool": "mgpu_prove_bench (single process, p2b_mgpu_*)", "n_gpus": G, "pool_retries": int(L.p2b_debug_pool_retries()), "shape": "%s 2^%d x %d wires" % (kind, n_log, num_wires),
                  "prove_ms": min(walls), "stages_ms": {k: min(v) for k, v in times.items()}, "wires_cap_word0": "%016x" % int(cap[0][0])}))
