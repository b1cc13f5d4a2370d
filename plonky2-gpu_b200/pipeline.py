"""The data path of prove() (plonky2/src/plonk/prover.rs:239-700 `my_prove`) end to end on the device, on a synthetic
witness: everything between "the witness matrix exists" and "the proof's field elements exist", i.e.

  commit wires -> Z / partial products -> commit them -> quotient values + coefficients -> commit the quotient chunks ->
  openings at zeta, g*zeta -> FRI opening proof (commit phase, PoW, query rounds)

with the transcript steps between the calls done by a stand-in (fixed challenges: the work does not depend on their values).
Not included (control plane, out of scope): circuit building, witness generation, the preprocessed constants_sigmas commit
(done once per circuit; built here at construction like the reference's CircuitData).

Shapes (BASELINE.json configs):
  "ecc"        config 4: wide_ecc_config (234 wires), the U32-heavy gate set of the ed25519-style circuit
               (ecdsa/src/gadgets/ecdsa.rs:64-110; gate list of cuda/plonky2_gpu_impl.cuh:600-685), ~2^17 rows
  "recursion"  config 3: standard_recursion_config (135 wires), the recursive verifier's gate set
               (plonky2/examples/bench_recursion.rs:175-207), 2^16 - 2^20 rows
Used by tools/prove_pipeline.py, bench.py --workload prove-* and tests/test_gpu_configs.py."""
import ctypes as C
import time

import numpy as np

from . import (GATE_ARITHMETIC, GATE_ARITHMETIC_EXTENSION, GATE_BASE_SUM, GATE_COMPARISON, GATE_CONSTANT, GATE_EXPONENTIATION,
               GATE_LOW_DEGREE_INTERPOLATION, GATE_MUL_EXTENSION, GATE_NOOP, GATE_POSEIDON, GATE_POSEIDON_MDS, GATE_PUBLIC_INPUT,
               GATE_RANDOM_ACCESS, GATE_REDUCING, GATE_REDUCING_EXTENSION, GATE_U32_ADD_MANY, GATE_U32_ARITHMETIC, GATE_U32_RANGE_CHECK,
               GATE_U32_SUBTRACTION, ORDER, Challenger, Circuit, DeviceBuffer, PolynomialBatch, _check, eval_openings, fri_prove_openings,
               lib, partial_products_and_zs)

SHAPES = {
    "ecc": dict(num_wires=234, num_routed=80, num_gate_consts=2,
                gates=[(GATE_NOOP, ()), (GATE_CONSTANT, (2,)), (GATE_PUBLIC_INPUT, ()), (GATE_ARITHMETIC, (20,)), (GATE_BASE_SUM, (63, 2)),
                       (GATE_BASE_SUM, (32, 2)), (GATE_RANDOM_ACCESS, (4, 4, 2)), (GATE_RANDOM_ACCESS, (2, 13, 2)), (GATE_U32_ARITHMETIC, (6,)),
                       (GATE_U32_ADD_MANY, (3, 9)), (GATE_U32_ADD_MANY, (5, 8)), (GATE_U32_RANGE_CHECK, (8,)), (GATE_U32_SUBTRACTION, (11,)),
                       (GATE_COMPARISON, (32, 16)), (GATE_COMPARISON, (8, 4)), (GATE_POSEIDON, ())],
                groups=[(0, 4), (4, 8), (8, 12), (12, 15), (15, 16)], sel=[0] * 4 + [1] * 4 + [2] * 4 + [3] * 3 + [4]),
    "recursion": dict(num_wires=135, num_routed=80, num_gate_consts=2,
                      gates=[(GATE_NOOP, ()), (GATE_CONSTANT, (2,)), (GATE_PUBLIC_INPUT, ()), (GATE_ARITHMETIC, (20,)),
                             (GATE_ARITHMETIC_EXTENSION, (10,)), (GATE_MUL_EXTENSION, (13,)), (GATE_REDUCING, (43,)),
                             (GATE_REDUCING_EXTENSION, (32,)), (GATE_BASE_SUM, (63, 2)), (GATE_RANDOM_ACCESS, (4, 4, 2)),
                             (GATE_EXPONENTIATION, (66,)), (GATE_POSEIDON_MDS, ()), (GATE_LOW_DEGREE_INTERPOLATION, (4,)), (GATE_POSEIDON, ())],
                      groups=[(0, 5), (5, 9), (9, 12), (12, 13), (13, 14)], sel=[0] * 5 + [1] * 4 + [2] * 3 + [3] + [4]),
}
STAGES = ("commit wires", "Z + partial products", "commit Z/pp", "quotient polys", "commit quotient chunks", "openings", "FRI prove_openings")


class ProvePipeline:
    rate_bits, cap_height, nc, qdf, pow_bits, queries = 3, 4, 2, 8, 16, 28

    def __init__(self, ctx, kind, n_log, seed=0):
        sh = SHAPES[kind]
        self.ctx, self.kind, self.n_log, self.n = ctx, kind, n_log, 1 << n_log
        self.num_wires, self.num_routed = sh["num_wires"], sh["num_routed"]
        self.gates, self.groups, self.sel = sh["gates"], sh["groups"], sh["sel"]
        self.num_constants = len(self.groups) + sh["num_gate_consts"]
        self.K = -(-self.num_routed // self.qdf)
        self.arity, d = [], n_log
        while d > 5 and d + self.rate_bits - 4 >= self.cap_height:   # ConstantArityBits(4, 5), fri/reduction_strategies.rs
            self.arity.append(4)
            d -= 4
        L = self.L = lib()
        self.k_is = [pow(7, j, ORDER) for j in range(self.num_routed)]
        self.circ = Circuit(self.gates, self.sel, self.groups, self.num_wires, self.num_routed, self.num_constants, self.k_is, n_log,
                            self.rate_bits, self.nc, self.qdf)
        rng = np.random.default_rng(seed)
        rnd = lambda k: [int(x) for x in rng.integers(0, ORDER, size=k, dtype=np.uint64)]   # noqa: E731
        self.pih, self.betas, self.gammas, self.alphas = rnd(4), rnd(self.nc), rnd(self.nc), rnd(self.nc)
        n = self.n
        # per-circuit data (outside any timed region): constants_sigmas values and their commitment
        self.d_cs = self._synth(self.num_constants + self.num_routed, 3)
        self.b_cs = PolynomialBatch.from_values(ctx, (self.d_cs, self.num_constants + self.num_routed, n), self.rate_bits, self.cap_height)
        self.d_sigma = DeviceBuffer(ctx, self.num_routed * n)   # sigma values = the last num_routed columns of constants_sigmas
        sig_host = np.empty(self.num_routed * n, dtype=np.uint64)
        L.p2b_memcpy_d2h(ctx.handle, sig_host.ctypes.data, C.c_void_p(self.d_cs.ptr + 8 * self.num_constants * n), 8 * self.num_routed * n)
        L.p2b_memcpy_h2d(ctx.handle, self.d_sigma.ptr, sig_host.ctypes.data, 8 * self.num_routed * n)
        self.d_wires = self._synth(self.num_wires, 1)            # the witness (random: every kernel's work is data-independent)
        ctx.synchronize()
        self.zeta = (0x123456789abcdef, 0xfedcba987654321)
        g = pow(1753635133440165772, 1 << (32 - n_log), ORDER)
        self.zeta_next = (self.zeta[0] * g % ORDER, self.zeta[1] * g % ORDER)
        self.size = self.circ.lde_size

    def _synth(self, cols, seed):
        dbuf = DeviceBuffer(self.ctx, cols * self.n)
        self.ctx.fill_synthetic(dbuf, cols * self.n, seed)
        return dbuf

    def describe(self):
        return "%s shape: 2^%d rows x %d wires, %d gates, rate %d, %d FRI reductions" % (self.kind, self.n_log, self.num_wires, len(self.gates),
                                                                                        self.rate_bits, len(self.arity))

    def _stage(self, name, fn, times):
        self.ctx.timer_start()
        r = fn()
        times.setdefault(name, []).append(self.ctx.timer_stop_ms())
        return r

    def prove(self, times=None, keep=False):
        """One pass of the data path.  Returns the host wall time in ms (and, with keep=True, the intermediate objects for
        parity checks: batches, quotient value / coefficient buffers, FRI proof -- the caller closes them)."""
        ctx, L, n = self.ctx, self.L, self.n
        times = {} if times is None else times
        arr = lambda x: (C.c_uint64 * len(x))(*x)   # noqa: E731
        t_all = time.perf_counter()
        b_w = self._stage("commit wires", lambda: PolynomialBatch.from_values(ctx, (self.d_wires, self.num_wires, n), self.rate_bits, self.cap_height), times)
        zs, shape = self._stage("Z + partial products", lambda: partial_products_and_zs(ctx, (self.d_wires, self.num_wires, n), (self.d_sigma, self.num_routed, n),
                                                                                       self.k_is, self.betas, self.gammas, self.qdf), times)
        b_z = self._stage("commit Z/pp", lambda: PolynomialBatch.from_values(ctx, (zs, shape[0], n), self.rate_bits, self.cap_height), times)
        dv, dc = DeviceBuffer(ctx, self.nc * self.size), DeviceBuffer(ctx, self.nc * self.size)
        self._stage("quotient polys", lambda: _check(L.p2b_quotient_polys(ctx.handle, C.byref(self.circ.struct), b_w.handle, b_z.handle, self.b_cs.handle,
                                                                         arr(self.pih), arr(self.betas), arr(self.gammas), arr(self.alphas), dv.ptr, dc.ptr)), times)
        # quotient_poly.chunks(degree) (prover.rs:151-166): [nc][8n] coefficients are already [nc*8][n] chunk-major
        b_q = self._stage("commit quotient chunks", lambda: PolynomialBatch.from_coeffs(ctx, (dc, self.nc * self.qdf, n), self.rate_bits, self.cap_height), times)
        oracles = [self.b_cs, b_w, b_z, b_q]

        def openings():
            out = [eval_openings(ctx, o, self.zeta) for o in oracles]
            out.append(eval_openings(ctx, b_z, self.zeta_next))
            return out
        opened = self._stage("openings", openings, times)
        polys = (self.num_constants + self.num_routed, self.num_wires, shape[0], self.nc * self.qdf)
        all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
        ch = Challenger(list(range(1, 13)), [5, 6, 7])
        pr = self._stage("FRI prove_openings", lambda: fri_prove_openings(ctx, oracles, [(self.zeta, all_polys), (self.zeta_next, [(2, p) for p in range(self.nc)])], ch,
                                                                         self.n_log, self.rate_bits, self.cap_height, self.pow_bits, self.queries, self.arity), times)
        ctx.synchronize()
        wall = (time.perf_counter() - t_all) * 1e3
        if keep:
            return wall, dict(b_w=b_w, b_z=b_z, b_q=b_q, quotient_values=dv, quotient_coeffs=dc, proof=pr, openings=opened, zs_shape=shape, zs=zs)
        pr.close()
        for b in (b_w, b_z, b_q):
            b.close()
        return wall

    def close(self):
        self.b_cs.close()
