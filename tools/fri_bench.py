"""FRI opening proof at the headline commit size: four committed batches shaped like the reference's oracles
(constants_sigmas 88, wires 135, zs_partial_products 20, quotient 16 polynomials of degree 2^n_log; standard_recursion_config
FRI: rate 3, cap 4, 16 PoW bits, 28 queries, arity 4 bits down to degree 2^5), all opened at zeta, Zs also at g*zeta.
Times p2b_eval_openings (OpeningSet::new) and p2b_fri_prove_openings end to end (host wall clock: the call ends with the
proof on the host) plus the device time from CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plonky2_gpu_b200 as p2b

n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
polys = (88, 135, 20, 16)
rate_bits, cap_height, pow_bits, queries = 3, 4, 16, 28
arity = []
d = n_log
while d > 5 and d + rate_bits - 4 >= cap_height:   # ConstantArityBits(4, 5), fri/reduction_strategies.rs:38-49
    arity.append(4)
    d -= 4
ctx = p2b.Context(0)
n = 1 << n_log
t0 = time.time()
oracles = []
for i, k in enumerate(polys):
    dbuf = p2b.DeviceBuffer(ctx, k * n)
    ctx.fill_synthetic(dbuf, k * n, 100 + i)
    oracles.append(p2b.PolynomialBatch.from_values(ctx, (dbuf, k, n), rate_bits, cap_height))
    del dbuf
ctx.synchronize()
print("committed %s polynomials of degree 2^%d in %.2f s" % (polys, n_log, time.time() - t0))
zeta = (0x123456789abcdef % p2b.ORDER, 0xfedcba987654321 % p2b.ORDER)
g = pow(1753635133440165772, 1 << (32 - n_log), p2b.ORDER)
zeta_next = (zeta[0] * g % p2b.ORDER, zeta[1] * g % p2b.ORDER)
all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
batches = [(zeta, all_polys), (zeta_next, [(2, p) for p in range(2)])]

def openings():
    for o in oracles:
        p2b.eval_openings(ctx, o, zeta)
    p2b.eval_openings(ctx, oracles[2], zeta_next)

def prove():
    ch = p2b.Challenger([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], [5, 6, 7])
    return p2b.fri_prove_openings(ctx, oracles, batches, ch, n_log, rate_bits, cap_height, pow_bits, queries, arity)

openings(); prove().close(); ctx.synchronize()
for name, fn in (("openings (5 eval_commitment calls)", openings), ("prove_openings", prove)):
    ts, ws = [], []
    for _ in range(5):
        ctx.synchronize()
        w = time.perf_counter()
        ctx.timer_start()
        r = fn()
        ts.append(ctx.timer_stop_ms())
        ws.append((time.perf_counter() - w) * 1e3)
        if r is not None:
            r.close()
    print("%s: device %.2f ms, wall %.2f ms (min of 5); arity bits %s" % (name, min(ts), min(ws), arity))
print("launches so far:", ctx.launch_count)
