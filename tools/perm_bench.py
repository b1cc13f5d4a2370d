"""Poseidon permutation / leaf-hash throughput of the library selected by P2B_LIB (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plonky2_gpu_b200 as p2b
from tests.golden.reference_kats import POSEIDON_KATS
ctx = p2b.Context(0); L = p2b.lib()
out = ctx.poseidon(np.array([k[0] for k in POSEIDON_KATS], dtype=np.uint64))
ok = all([int(x) for x in r] == e for r, (_, e) in zip(out, POSEIDON_KATS))
rng = np.random.default_rng(1)
N, P = 1 << 20, 135
leaves = p2b.DeviceBuffer(ctx, N * P); ctx.fill_synthetic(leaves, N * P, 7)
dig = p2b.DeviceBuffer(ctx, 2 * N * 4); cap = p2b.DeviceBuffer(ctx, 64)
def merkle(): p2b._check(L.p2b_merkle_tree(ctx.handle, leaves.ptr, N, P, P, 1, 4, dig.ptr, cap.ptr))
merkle(); ctx.synchronize()
ts = []
for _ in range(3):
    ctx.timer_start(); merkle(); ts.append(ctx.timer_stop_ms())
perms = N * 17 + N - 16
capv = cap.to_host(64)
print("%s  KAT %s  merkle 2^20x135: %.2f ms  -> %.1f Mperm/s  cap0 %016x" % (os.environ.get("P2B_LIB", "default"), "ok" if ok else "FAIL", min(ts), perms / min(ts) / 1e3, int(capv[0])))
