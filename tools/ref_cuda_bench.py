"""GPU-vs-GPU baseline ("the kernel to beat on the same box"): the reference's own CUDA translation
unit (cuda/plonky2_gpu.cu, compiled UNMODIFIED for sm_100a into oracle/_ref/libplonky2_ref_cuda.so by oracle/Makefile)
timed on the headline commit -- `ifft` (plonky2_gpu.cu:70-86) + `merkle_tree_from_coeffs` (:435-606) on 2^n_log x P,
rate_bits 3, cap_height 4 -- with CUDA events, beside this library's drop-in symbols of the same name on the same
buffers, and the caps of both compared.

    python tools/ref_cuda_bench.py [n_log=20] [P=135] [reps=3]

Prints one JSON line.  Test/measurement infrastructure: nothing in the product imports this.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libplonky2_ref_cuda.so")
SEED = 0x504C4F4E4B5932


class RefStreams(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("stream2", C.c_void_p)]


def measure(n_log=20, P=135, reps=3, rate_bits=3, cap_height=4, ctx=None, with_ours=True):
    """Returns a dict: ref_ms (ifft + merkle_tree_from_coeffs of the reference's kernels), ours_compat_ms (the drop-in
    symbols of this library on the same layout), caps_equal."""
    import torch
    import oracle
    import plonky2_gpu_b200 as p2b
    if not os.path.exists(REF_SO):
        return {"unavailable": "oracle/_ref/libplonky2_ref_cuda.so not built (needs /root/reference at build time)"}
    own_ctx = ctx is None
    if own_ctx:
        p2b.build()
        ctx = p2b.Context(0)
    L = p2b.lib()
    ref = C.CDLL(REF_SO)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    streams = RefStreams(s1.cuda_stream, s2.cuda_stream)
    n = 1 << n_log
    N = n << rate_bits
    ncap = 1 << cap_height
    nd = 2 * (N - ncap)
    pad = N * P
    assert pad < 2 ** 31, "the reference ABI passes pad_extvalues_len as a C int"
    total = 2 * pad + 4 * (nd + ncap)
    base = p2b.DeviceBuffer(ctx, total)
    vals = p2b.DeviceBuffer(ctx, P * n)
    ctx.fill_synthetic(vals, P * n, SEED)
    root1 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log))
    root2 = p2b.DeviceBuffer.from_host(ctx, oracle.fft_root_table_concat(n_log + rate_bits))
    sp = np.empty(n, dtype=np.uint64)
    cur = 1
    for i in range(n):
        sp[i] = cur
        cur = cur * 7 % oracle.ORDER
    shift = p2b.DeviceBuffer.from_host(ctx, sp)
    n_inv = C.c_uint64(oracle.inverse_2exp(n_log))
    ctx.synchronize()

    def run(lib):
        lib.ifft.restype = p2b.RustError
        lib.merkle_tree_from_coeffs.restype = p2b.RustError
        times = []
        cap = None
        for _ in range(reps + 1):
            C.cdll.LoadLibrary('libcudart.so.12').cudaMemcpy(C.c_void_p(base.ptr), C.c_void_p(vals.ptr), C.c_size_t(P * n * 8), C.c_int(3))
            ctx.synchronize()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s1):
                e0.record(s1)
                e = lib.ifft(C.c_void_p(base.ptr), C.c_int(P), C.c_int(n), C.c_int(n_log), C.c_void_p(root1.ptr), C.byref(n_inv),
                             C.byref(streams))
                assert e.code == 0
                e = lib.merkle_tree_from_coeffs(C.c_void_p(base.ptr), C.c_void_p(base.ptr), C.c_int(P), C.c_int(n), C.c_int(n_log),
                                                C.c_void_p(root1.ptr), C.c_void_p(root2.ptr), C.c_void_p(shift.ptr), C.c_int(rate_bits),
                                                C.c_int(0), C.c_int(cap_height), C.c_int(pad), C.byref(streams))
                assert e.code == 0
                s1.wait_stream(s2)
                e1.record(s1)
            torch.cuda.synchronize()
            ctx.synchronize()
            times.append(e0.elapsed_time(e1))
            cap = base.to_host(4 * ncap, offset=2 * pad + 4 * nd).reshape(ncap, 4).copy()
            cap[cap >= np.uint64(oracle.ORDER)] -= np.uint64(oracle.ORDER)
        return min(times[1:]), cap

    out = {"workload": "ifft + merkle_tree_from_coeffs 2^%d x %d, rate_bits %d, cap_height %d (reference device layout)" % (n_log, P, rate_bits, cap_height)}
    ref_ms, ref_cap = run(ref)
    out["ref_ms"] = ref_ms
    out["ref_cap_word0"] = "%016x" % int(ref_cap[0][0])
    if with_ours:
        ours_ms, ours_cap = run(L)
        out["ours_compat_ms"] = ours_ms
        out["caps_equal"] = bool(np.array_equal(ref_cap, ours_cap))
    if own_ctx:
        ctx.close()
    return out


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    print(json.dumps(measure(*a)))
