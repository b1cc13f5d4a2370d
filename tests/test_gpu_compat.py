"""GPU: the reference's legacy C symbols (cuda/src/lib.rs:52-145) beyond what test_ref_cuda_crosscheck.py covers: blinding
(salt_size = 4) through merkle_tree_from_coeffs, fft_blinding (plonky2_gpu.cu:88-136), and two host threads calling the
legacy entry points at once (the shared compat context is locked for the whole call)."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
P_ = oracle.ORDER


class RefStreams(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("stream2", C.c_void_p)]


@pytest.fixture(scope="module")
def ctx():
    p2b.build()
    c = p2b.Context(0)
    yield c
    c.close()


def _bitrev_perm(log_n):
    n = 1 << log_n
    return np.array([oracle.reverse_bits(i, log_n) for i in range(n)], dtype=np.int64)


def test_merkle_tree_from_coeffs_with_salt_columns(ctx):
    """oracle.rs:302 passes salt_size = 4 for blinding oracles; the reference hashes whatever columns [P, P+4) of the work
    area hold, indexed by leaf (plonky2_gpu.cu:552-600).  With known words there the result must be the CPU path's."""
    rng = np.random.default_rng(11)
    n_log, Pn, rate_bits, cap_height, salt_size = 9, 7, 3, 2, 4
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    coeffs = rng.integers(0, P_, size=(Pn, n), dtype=np.uint64)
    salt = rng.integers(0, P_, size=(salt_size, N), dtype=np.uint64)           # natural point order (oracle.rs:998-1002)
    exp = oracle.batch_from_coeffs(coeffs, rate_bits, cap_height, salt=salt)
    ncap, nd = 1 << cap_height, 2 * (N - (1 << cap_height))
    leaf_len = Pn + salt_size
    pad = N * leaf_len
    base = p2b.DeviceBuffer(ctx, 2 * pad + 4 * (nd + ncap))
    L = p2b.lib()
    p2b._check(L.p2b_memcpy_h2d(ctx.handle, base.ptr, coeffs.ctypes.data, coeffs.size * 8))
    salt_by_leaf = np.ascontiguousarray(salt[:, _bitrev_perm(n_log + rate_bits)])  # leaf L <- salt[:, reverse_bits(L)]
    p2b._check(L.p2b_memcpy_h2d(ctx.handle, base.ptr + 8 * (pad + Pn * N), salt_by_leaf.ctypes.data, salt_by_leaf.size * 8))
    L.merkle_tree_from_coeffs.restype = p2b.RustError
    e = L.merkle_tree_from_coeffs(C.c_void_p(base.ptr), C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), None, None, None,
                                  C.c_int(rate_bits), C.c_int(salt_size), C.c_int(cap_height), C.c_int(pad), None)
    assert e.code == 0, e.message
    ctx.synchronize()
    assert np.array_equal(base.to_host(N * leaf_len).reshape(N, leaf_len), exp.leaves)
    assert np.array_equal(base.to_host(4 * nd, offset=2 * pad).reshape(nd, 4), exp.digests)
    assert np.array_equal(base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4), exp.cap)


@pytest.mark.parametrize("n_log,Pn,rate_bits", [(8, 5, 3), (10, 3, 1), (6, 2, 2)])
def test_fft_blinding_natural_order_lde(ctx, n_log, Pn, rate_bits):
    rng = np.random.default_rng(3 + n_log)
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    coeffs = rng.integers(0, P_, size=(Pn, n), dtype=np.uint64)
    pad = N * Pn
    base = p2b.DeviceBuffer(ctx, 2 * pad)
    L = p2b.lib()
    p2b._check(L.p2b_memcpy_h2d(ctx.handle, base.ptr, coeffs.ctypes.data, coeffs.size * 8))
    L.fft_blinding.restype = p2b.RustError
    e = L.fft_blinding(C.c_void_p(base.ptr), C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), None, None, C.c_int(rate_bits),
                       C.c_int(pad), None)
    assert e.code == 0, e.message
    got = base.to_host(N * Pn, offset=pad).reshape(Pn, N)
    want = np.stack([oracle.lde_coset_fft(coeffs[c], rate_bits) for c in range(Pn)])
    assert np.array_equal(got, want)
    assert np.array_equal(base.to_host(n * Pn).reshape(Pn, n), coeffs)   # the coefficients are left in place


def test_legacy_symbols_from_two_host_threads(ctx):
    """compat.cuh serialises legacy calls on the shared context: concurrent `ifft`s must both be right."""
    L = p2b.lib()
    L.ifft.restype = p2b.RustError
    rng = np.random.default_rng(77)
    n_log, Pn = 12, 9
    n = 1 << n_log
    vals = [rng.integers(0, P_, size=(Pn, n), dtype=np.uint64) for _ in range(2)]
    bufs = [p2b.DeviceBuffer.from_host(ctx, v.reshape(-1)) for v in vals]
    want = [np.stack([oracle.ifft(v[c]) for c in range(Pn)]) for v in vals]
    errs = [None, None]

    def work(k):
        for _ in range(20):
            p2b._check(L.p2b_memcpy_h2d(ctx.handle, bufs[k].ptr, vals[k].ctypes.data, vals[k].size * 8))
            e = L.ifft(C.c_void_p(bufs[k].ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), None, None, None)
            if e.code != 0 or not np.array_equal(bufs[k].to_host(Pn * n).reshape(Pn, n), want[k]):
                errs[k] = "mismatch"
                return

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert errs == [None, None]


def test_stage_trace_names_follow_the_reference_timing_tree(ctx):
    """p2b_ctx_trace: stages carry the reference's `timed!` scope names (fri/oracle.rs:717-966)."""
    rng = np.random.default_rng(2)
    values = rng.integers(0, P_, size=(20, 1 << 10), dtype=np.uint64)
    ctx.trace(True)
    b = p2b.PolynomialBatch.from_values(ctx, values, 3, 4)
    rep = dict((name, (calls, ms)) for name, calls, ms in ctx.trace_report())
    ctx.trace(False)
    b.close()
    for name in ("IFFT", "FFT + blinding", "build Merkle tree: leaf hashes", "build Merkle tree: digest layers"):
        assert name in rep and rep[name][0] >= 1 and rep[name][1] > 0.0
    assert rep["FFT + blinding"][0] % 8 == 0        # one LDE per coset block (and per column group of the pipelined upload)
