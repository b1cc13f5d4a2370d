"""GPU: cross-check against the REFERENCE ITSELF.  oracle/_ref/libplonky2_ref_cuda.so is the reference's own CUDA
translation unit (cuda/plonky2_gpu.cu) compiled unmodified for sm_100a by oracle/Makefile (`make ref`, in the build
container where /root/reference exists; the .so travels to the GPU box).  Its `ifft` and `merkle_tree_from_coeffs` are
run on the same inputs as (i) the CPU oracle and (ii) this library's drop-in symbols of the same name, with the
reference's device-memory layout (cuda/plonky2_gpu.cu:435-606, fri/oracle.rs:316-455).  This pins the oracle's LDE
values, digests and caps to outputs of the reference, not only to its Poseidon/bit-reversal known answers."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libplonky2_ref_cuda.so")
P_ = oracle.ORDER


class RefStreams(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("stream2", C.c_void_p)]


def canon(a):
    a = a.copy()
    a[a >= np.uint64(P_)] -= np.uint64(P_)
    return a


@pytest.fixture(scope="module")
def env():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libplonky2_ref_cuda.so not built (needs /root/reference at build time)")
    import torch
    p2b.build()
    ctx = p2b.Context(0)
    ref = C.CDLL(REF_SO)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    streams = RefStreams(s1.cuda_stream, s2.cuda_stream)
    yield ctx, ref, streams, (s1, s2)
    ctx.close()


def run_layout(fn_ifft, fn_tree, ctx, streams, values, rate_bits, cap_height):
    """Drive `ifft` + `merkle_tree_from_coeffs` the way fri/oracle.rs:352-457 does; returns coeffs, leaves, digests, cap."""
    Pn, n = values.shape
    n_log = n.bit_length() - 1
    N = n << rate_bits
    ncap = 1 << cap_height
    nd = 2 * (N - ncap)
    pad = N * Pn
    total = 2 * pad + 4 * (nd + ncap)
    base = p2b.DeviceBuffer(ctx, total)
    p2b._check(p2b.lib().p2b_memcpy_h2d(ctx.handle, base.ptr, values.ctypes.data, values.size * 8))
    root1 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log))
    root2 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log + rate_bits))
    sp = np.empty(n, dtype=np.uint64)
    cur = 1
    for i in range(n):
        sp[i] = cur
        cur = cur * 7 % P_
    shift = p2b.DeviceBuffer.from_host(ctx, sp)
    n_inv = C.c_uint64(oracle.inverse_2exp(n_log))
    fn_ifft.restype = p2b.RustError
    fn_tree.restype = p2b.RustError
    e = fn_ifft(C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), C.c_void_p(root1.ptr), C.byref(n_inv), C.byref(streams))
    assert e.code == 0
    coeffs = base.to_host(Pn * n).reshape(Pn, n)
    e = fn_tree(C.c_void_p(base.ptr), C.c_void_p(base.ptr), C.c_int(Pn), C.c_int(n), C.c_int(n_log), C.c_void_p(root1.ptr),
                C.c_void_p(root2.ptr), C.c_void_p(shift.ptr), C.c_int(rate_bits), C.c_int(0), C.c_int(cap_height), C.c_int(pad),
                C.byref(streams))
    assert e.code == 0
    ctx.synchronize()
    leaves = base.to_host(N * Pn).reshape(N, Pn)
    dig = base.to_host(4 * nd, offset=2 * pad).reshape(nd, 4)
    cap = base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4)
    return canon(coeffs), canon(leaves), canon(dig), canon(cap)


@pytest.mark.parametrize("n_log,Pn,cap_height", [(9, 5, 4), (10, 7, 4), (12, 20, 4), (11, 9, 0), (13, 135, 4)])
def test_reference_cuda_vs_oracle_vs_ours(env, n_log, Pn, cap_height):
    ctx, ref, streams, _ = env
    rng = np.random.default_rng(4242 + n_log)
    values = rng.integers(0, P_, size=(Pn, 1 << n_log), dtype=np.uint64)
    # Pn >= 5: the reference CUDA hashes leaves of <= 4 elements too, unlike its CPU path (hash_or_noop, config.rs:56-67)
    rate_bits = 3  # the reference kernels hard-code 2^3 - 1 zero blocks (plonky2_gpu_impl.cuh:290-294)
    r_coeffs, r_leaves, r_dig, r_cap = run_layout(ref.ifft, ref.merkle_tree_from_coeffs, ctx, streams, values, rate_bits, cap_height)
    L = p2b.lib()
    o_coeffs, o_leaves, o_dig, o_cap = run_layout(L.ifft, L.merkle_tree_from_coeffs, ctx, streams, values, rate_bits, cap_height)
    e = oracle.batch_from_values(values, rate_bits, cap_height)
    # reference GPU == oracle (pins the oracle to the reference's own outputs)
    assert np.array_equal(r_coeffs, e.coeffs)
    assert np.array_equal(r_leaves, e.leaves)
    assert np.array_equal(r_cap, e.cap)
    assert np.array_equal(r_dig, e.digests)
    # this library's drop-in symbols == reference GPU, in the reference's memory layout
    assert np.array_equal(o_coeffs, r_coeffs)
    assert np.array_equal(o_leaves, r_leaves)
    assert np.array_equal(o_dig, r_dig)
    assert np.array_equal(o_cap, r_cap)


def test_compat_build_merkle_tree_and_transpose(env):
    """build_merkle_tree (lib.rs:71-81) + transpose (plonky2_gpu.cu:192-215) on a column-major LDE in the work area."""
    ctx, ref, streams, _ = env
    rng = np.random.default_rng(7)
    n_log, Pn, rate_bits, cap_height = 8, 6, 2, 3
    n, N = 1 << n_log, 1 << (n_log + rate_bits)
    coeffs = rng.integers(0, P_, size=(Pn, n), dtype=np.uint64)
    lde_cols = np.stack([oracle.lde_coset_fft(coeffs[c], rate_bits) for c in range(Pn)])  # natural order, column-major
    pad = N * Pn
    ncap, nd = 1 << cap_height, 2 * (N - (1 << cap_height))
    base = p2b.DeviceBuffer(ctx, 2 * pad + 4 * (nd + ncap))
    p2b._check(p2b.lib().p2b_memcpy_h2d(ctx.handle, base.ptr + 8 * pad, lde_cols.ctypes.data, lde_cols.size * 8))
    L = p2b.lib()
    e = L.build_merkle_tree(base.ptr, Pn, n, n_log, rate_bits, 0, cap_height, pad, C.byref(streams))
    assert e.code == 0
    e = L.transpose(base.ptr, Pn, n, rate_bits, 0, pad, C.byref(streams))
    assert e.code == 0
    exp = oracle.batch_from_coeffs(coeffs, rate_bits, cap_height)
    assert np.array_equal(base.to_host(N * Pn).reshape(N, Pn), exp.leaves)
    assert np.array_equal(base.to_host(4 * nd, offset=2 * pad).reshape(nd, 4), exp.digests)
    assert np.array_equal(base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4), exp.cap)
