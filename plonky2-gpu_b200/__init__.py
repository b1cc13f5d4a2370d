"""plonky2_gpu_b200 -- host-side Python binding of libplonky2_b200.so (ctypes over the C ABI in
include/plonky2_b200.h).  Used by tests/, bench.py and __graft_entry__.py; mirrors the reference's
PolynomialBatch / MerkleTree interface (plonky2/src/fri/oracle.rs:112-120, 709-1018;
plonky2/src/hash/merkle_tree.rs:41-66, 383-440) so the parity tests read like the reference's own.

There is NO CPU fallback: loading fails loudly if the CUDA library is missing, and every call fails if no
sm_100 device is present.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2B_LIB") or os.path.join(_HERE, "libplonky2_b200.so")  # P2B_LIB: experimental variants only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "plonky2_b200.h")
ORDER = 0xFFFFFFFF00000001
SALT_SIZE = 4

P2B_OK, P2B_ERR_INVALID, P2B_ERR_CUDA, P2B_ERR_OOM, P2B_ERR_UNSUPPORTED = 0, 1, 2, 3, 4


class P2BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("p2b error %d: %s" % (code, msg))
        self.code = code


def _load_build_module():
    spec = importlib.util.spec_from_file_location("p2b_build", os.path.join(_HERE, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if os.environ.get("P2B_LIB"):
        return LIB_PATH
    return _load_build_module().build(force=force, verbose=verbose)


class BatchInfo(C.Structure):
    _fields_ = [("degree_log", C.c_uint32), ("rate_bits", C.c_uint32), ("cap_height", C.c_uint32),
                ("salt_size", C.c_uint32), ("num_polys", C.c_uint64), ("num_leaves", C.c_uint64),
                ("leaf_len", C.c_uint64), ("num_digests", C.c_uint64)]


class DataSlice(C.Structure):
    """DataSlice (cuda/src/lib.rs:52-56)"""
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_int)]


class Gate(C.Structure):
    """p2b_gate (include/plonky2_b200.h): one entry of common_data.gates with its selector group."""
    _fields_ = [("type", C.c_uint32), ("selector_index", C.c_uint32), ("group_start", C.c_uint32), ("group_end", C.c_uint32),
                ("p0", C.c_uint32), ("p1", C.c_uint32), ("p2", C.c_uint32), ("reserved", C.c_uint32)]


class CircuitStruct(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("rate_bits", C.c_uint32), ("quotient_degree_factor", C.c_uint32),
                ("num_challenges", C.c_uint32), ("num_wires", C.c_uint32), ("num_routed_wires", C.c_uint32),
                ("num_constants", C.c_uint32), ("num_selectors", C.c_uint32), ("num_gates", C.c_uint32), ("reserved", C.c_uint32),
                ("gates", C.POINTER(Gate)), ("k_is", C.POINTER(C.c_uint64))]


class ChallengerStruct(C.Structure):
    """p2b_challenger == Challenger {sponge_state, input_buffer, output_buffer} (iop/challenger.rs:15-21)."""
    _fields_ = [("sponge_state", C.c_uint64 * 12), ("input_buffer", C.c_uint64 * 8), ("output_buffer", C.c_uint64 * 8),
                ("input_len", C.c_uint32), ("output_len", C.c_uint32)]


class FriPolyInfo(C.Structure):
    _fields_ = [("oracle_index", C.c_uint32), ("polynomial_index", C.c_uint32)]


class FriBatchInfoStruct(C.Structure):
    _fields_ = [("point", C.c_uint64 * 2), ("polynomials", C.POINTER(FriPolyInfo)), ("num_polynomials", C.c_uint32),
                ("reserved", C.c_uint32)]


class FriParamsStruct(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("rate_bits", C.c_uint32), ("cap_height", C.c_uint32),
                ("proof_of_work_bits", C.c_uint32), ("num_query_rounds", C.c_uint32), ("num_reductions", C.c_uint32),
                ("reduction_arity_bits", C.POINTER(C.c_uint32))]


class FriProofInfo(C.Structure):
    _fields_ = [("num_reductions", C.c_uint32), ("num_query_rounds", C.c_uint32), ("num_oracles", C.c_uint32),
                ("cap_height", C.c_uint32), ("final_poly_len", C.c_uint64), ("lde_size", C.c_uint64)]


GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_BASE_SUM, GATE_POSEIDON, GATE_RANDOM_ACCESS = range(7)
GATE_U32_ARITHMETIC, GATE_U32_ADD_MANY, GATE_U32_RANGE_CHECK, GATE_U32_SUBTRACTION, GATE_COMPARISON = range(7, 12)
(GATE_ARITHMETIC_EXTENSION, GATE_MUL_EXTENSION, GATE_REDUCING, GATE_REDUCING_EXTENSION, GATE_EXPONENTIATION, GATE_POSEIDON_MDS,
 GATE_HIGH_DEGREE_INTERPOLATION, GATE_LOW_DEGREE_INTERPOLATION) = range(12, 20)


class Circuit:
    """The part of CommonCircuitData that compute_quotient_polys reads (plonk/circuit_data.rs:270-349).
    gates: list of (type, params tuple); selector_indices[i], groups[selector] = (start, end) as SelectorsInfo."""

    def __init__(self, gates, selector_indices, groups, num_wires, num_routed_wires, num_constants, k_is, degree_bits,
                 rate_bits=3, num_challenges=2, quotient_degree_factor=8):
        self._gates = (Gate * max(len(gates), 1))()
        for i, (t, params) in enumerate(gates):
            pr = list(params) + [0, 0, 0]
            g0, g1 = groups[selector_indices[i]]
            self._gates[i] = Gate(t, selector_indices[i], g0, g1, pr[0], pr[1], pr[2], 0)
        self._kis = (C.c_uint64 * max(num_routed_wires, 1))(*[int(k) % ORDER for k in k_is])
        self.struct = CircuitStruct(degree_bits, rate_bits, quotient_degree_factor, num_challenges, num_wires, num_routed_wires,
                                    num_constants, len(groups), len(gates), 0, self._gates, self._kis)
        self.num_challenges, self.degree_bits = num_challenges, degree_bits
        self.quotient_degree_bits = (quotient_degree_factor - 1).bit_length()

    @property
    def lde_size(self):
        return 1 << (self.degree_bits + self.quotient_degree_bits)


class RustError(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_void_p)]


_lib = None


def lib():
    """The loaded shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python plonky2-gpu_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    sigs = {
        "p2b_last_error": (C.c_char_p, []),
        "p2b_version": (C.c_char_p, []),
        "p2b_debug_pool_retries": (C.c_ulonglong, []),
        "p2b_ctx_create": (i, [i, C.POINTER(vp)]),
        "p2b_ctx_destroy": (None, [vp]),
        "p2b_ctx_stream": (vp, [vp]),
        "p2b_ctx_synchronize": (i, [vp]),
        "p2b_ctx_launch_count": (u64, [vp]),
        "p2b_ctx_time_leaf_hash": (i, [vp, i]),
        "p2b_ctx_trace": (i, [vp, i]),
        "p2b_ctx_trace_report": (i, [vp, C.c_char_p, u64]),
        "p2b_ctx_debug_force_exact_redo": (i, [vp, i]),
        "p2b_ctx_leaf_hash_time": (i, [vp, C.POINTER(C.c_double), C.POINTER(u64)]),
        "p2b_commit_from_values": (i, [vp, vp, i, u32, u64, u32, u32, vp, i, C.POINTER(vp)]),
        "p2b_commit_from_coeffs": (i, [vp, vp, i, u32, u64, u32, u32, vp, i, C.POINTER(vp)]),
        "p2b_commit_from_values_ex": (i, [vp, vp, i, u32, u64, u32, u32, vp, i, vp, C.POINTER(vp)]),
        "p2b_batch_destroy": (None, [vp]),
        "p2b_batch_get_info": (i, [vp, C.POINTER(BatchInfo)]),
        "p2b_batch_device_ptrs": (i, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "p2b_batch_get_coeffs": (i, [vp, vp]),
        "p2b_batch_get_cap": (i, [vp, vp]),
        "p2b_batch_get_digests": (i, [vp, vp]),
        "p2b_batch_get_leaves": (i, [vp, u64, u64, vp]),
        "p2b_batch_get_lde_values": (i, [vp, u64, u64, vp]),
        "p2b_batch_prove": (i, [vp, vp, u64, vp]),
        "p2b_batch_open_rows": (i, [vp, vp, u64, vp, vp]),
        "p2b_commit_blocks": (i, [vp, vp, u32, u64, u32, u32, vp, u64, u64, C.POINTER(vp)]),
        "p2b_commit_blocks_begin": (i, [vp, u32, u64, u32, u32, u64, u64, C.POINTER(vp)]),
        "p2b_commit_blocks_absorb": (i, [vp, vp, u64, u64]),
        "p2b_commit_blocks_finish": (i, [vp]),
        "p2b_batch_shard_info": (i, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]),
        "p2b_batch_export_nodes": (i, [vp, u32, u64, u64, vp]),
        "p2b_batch_import_nodes": (i, [vp, u32, u64, u64, vp]),
        "p2b_batch_finish_layers": (i, [vp, u32]),
        "p2b_quotient_polys": (i, [vp, C.POINTER(CircuitStruct), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "p2b_quotient_polys_rows": (i, [vp, C.POINTER(CircuitStruct), vp, u64, vp, u64, vp, u64, vp, vp, vp, vp, vp, vp]),
        "p2b_partial_products_and_zs": (i, [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp]),
        "p2b_eval_openings": (i, [vp, vp, vp, vp]),
        "p2b_fri_prove_openings": (i, [vp, vp, u32, C.POINTER(FriBatchInfoStruct), u32, C.POINTER(ChallengerStruct),
                                       C.POINTER(FriParamsStruct), C.POINTER(vp)]),
        "p2b_fri_proof_destroy": (None, [vp]),
        "p2b_fri_proof_get_info": (i, [vp, C.POINTER(FriProofInfo)]),
        "p2b_fri_proof_get_cap": (i, [vp, u32, vp]),
        "p2b_fri_proof_get_final_poly": (i, [vp, vp]),
        "p2b_fri_proof_get_pow_witness": (i, [vp, vp]),
        "p2b_fri_proof_get_query_indices": (i, [vp, vp]),
        "p2b_fri_proof_get_initial": (i, [vp, u32, vp, vp]),
        "p2b_fri_proof_get_step": (i, [vp, u32, vp, vp, C.POINTER(u32)]),
        "p2b_fri_proof_get_debug": (i, [vp, u32, vp]),
        "p2b_ifft_batch": (i, [vp, vp, vp, u32, u64]),
        "p2b_lde_leaves": (i, [vp, vp, u32, u64, u32, vp, u64, u64]),
        "p2b_merkle_tree": (i, [vp, vp, u64, u64, u64, u64, u32, vp, vp]),
        "p2b_poseidon_permute": (i, [vp, vp, u64]),
        "p2b_field_op": (i, [vp, i, vp, vp, vp, u64]),
        "p2b_fill_synthetic": (i, [vp, vp, u64, u64, u64]),
        "p2b_malloc": (i, [vp, u64, C.POINTER(vp)]),
        "p2b_free": (i, [vp, vp]),
        "p2b_malloc_host": (i, [u64, C.POINTER(vp)]),
        "p2b_free_host": (i, [vp]),
        "p2b_memcpy_h2d": (i, [vp, vp, vp, u64]),
        "p2b_memcpy_d2h": (i, [vp, vp, vp, u64]),
        "p2b_timer_start": (i, [vp]),
        "p2b_timer_stop_ms": (i, [vp, C.POINTER(C.c_float)]),
        # single-process multi-device commit
        "p2b_mgpu_create": (i, [C.POINTER(C.c_int), i, C.POINTER(vp)]),
        "p2b_mgpu_destroy": (None, [vp]),
        "p2b_mgpu_device_count": (i, [vp]),
        "p2b_mgpu_ctx": (vp, [vp, i]),
        "p2b_mgpu_peer_access": (i, [vp]),
        "p2b_mgpu_synchronize": (i, [vp]),
        "p2b_mgpu_timer_start": (i, [vp]),
        "p2b_mgpu_timer_stop_ms": (i, [vp, C.POINTER(C.c_float)]),
        "p2b_mgpu_commit_from_values": (i, [vp, vp, u32, u64, u32, u32, vp, C.POINTER(vp)]),
        "p2b_mgpu_resident_cols": (i, [vp, i, u32, u64, C.POINTER(vp), C.POINTER(u64)]),
        "p2b_mgpu_commit_resident": (i, [vp, u32, u64, u32, u32, vp, C.POINTER(vp)]),
        "p2b_mgpu_round": (i, [vp, u64, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "p2b_mgpu_batch_destroy": (None, [vp]),
        "p2b_mgpu_batch_get_info": (i, [vp, C.POINTER(BatchInfo)]),
        "p2b_mgpu_batch_shard": (vp, [vp, i]),
        "p2b_mgpu_batch_get_cap": (i, [vp, vp]),
        "p2b_mgpu_batch_open_rows": (i, [vp, vp, u64, vp, vp]),
        "p2b_mgpu_batch_get_leaves": (i, [vp, u64, u64, vp]),
        "p2b_mgpu_commit_from_device_coeffs": (i, [vp, C.POINTER(vp), u32, u64, u32, u32, C.POINTER(vp)]),
        "p2b_mgpu_commit_from_device_values": (i, [vp, i, vp, u32, u64, u32, u32, C.POINTER(vp)]),
        "p2b_mgpu_quotient_polys": (i, [vp, C.POINTER(CircuitStruct), vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "p2b_mgpu_eval_openings": (i, [vp, vp, vp, vp]),
        "p2b_mgpu_fri_prove_openings": (i, [vp, C.POINTER(vp), u32, C.POINTER(FriBatchInfoStruct), u32, C.POINTER(ChallengerStruct),
                                            C.POINTER(FriParamsStruct), C.POINTER(vp)]),
        "fft_blinding": (RustError, [vp, vp, i, i, i, vp, vp, i, i, vp]),
        # reference-compatible symbols (cuda/src/lib.rs:52-145)
        "init": (None, []),
        "ifft": (RustError, [vp, i, i, i, vp, vp, vp]),
        "build_merkle_tree": (RustError, [vp, i, i, i, i, i, i, i, vp]),
        "merkle_tree_from_values": (RustError, [vp, vp, i, i, i, vp, vp, vp, vp, i, i, i, i, vp]),
        "merkle_tree_from_coeffs": (RustError, [vp, vp, i, i, i, vp, vp, vp, i, i, i, i, vp]),
        "transpose": (RustError, [vp, i, i, i, i, i, vp]),
        "compute_quotient_polys": (RustError, [vp, i, i, i, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "p2b_compat_set_circuit": (i, [C.POINTER(CircuitStruct), vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _check(rc):
    if rc != P2B_OK:
        raise P2BError(rc, lib().p2b_last_error().decode())


def _np(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


class DeviceBuffer:
    """A device allocation of u64 elements owned by a Context."""

    def __init__(self, ctx, count):
        self.ctx, self.count = ctx, int(count)
        p = C.c_void_p()
        _check(lib().p2b_malloc(ctx.handle, self.count * 8, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, ctx, arr):
        arr = _np(arr)
        b = cls(ctx, arr.size)
        if arr.size:
            _check(lib().p2b_memcpy_h2d(ctx.handle, b.ptr, arr.ctypes.data, arr.size * 8))
        return b

    def to_host(self, count=None, offset=0):
        count = self.count - offset if count is None else count
        out = np.empty(count, dtype=np.uint64)
        if count:
            _check(lib().p2b_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr + 8 * offset, count * 8))
        return out

    def free(self):
        if self.ptr:
            lib().p2b_free(self.ctx.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint64 array (for the end-to-end benchmark leg)."""

    def __init__(self, count):
        p = C.c_void_p()
        _check(lib().p2b_malloc_host(int(count) * 8, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint64)), shape=(int(count),))

    def free(self):
        if self.ptr:
            self.array = None
            lib().p2b_free_host(self.ptr)
            self.ptr = None


class Context:
    """Per-device context (replaces the reference's CudaInvContext, fri/oracle.rs:75-109)."""

    def __init__(self, device=-1, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:          # a context owned by a MultiGpu group (MultiGpu.context)
            self.handle = _borrowed
            return
        h = C.c_void_p()
        _check(lib().p2b_ctx_create(device, C.byref(h)))
        self.handle = h.value

    def close(self):
        if self.handle and self._owned:
            lib().p2b_ctx_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib().p2b_ctx_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(lib().p2b_ctx_launch_count(self.handle))

    def debug_force_exact_redo(self, enable=True):
        _check(lib().p2b_ctx_debug_force_exact_redo(self.handle, 1 if enable else 0))

    def trace(self, enable=True):
        """Per-stage device timing under the reference's TimingTree scope names (util/timing.rs:8-192)."""
        _check(lib().p2b_ctx_trace(self.handle, 1 if enable else 0))

    def trace_report(self):
        """[(stage name, calls, total ms)] since tracing was enabled / last reported."""
        buf = C.create_string_buffer(1 << 16)
        _check(lib().p2b_ctx_trace_report(self.handle, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, calls, ms = line.split("\t")
            out.append((name, int(calls), float(ms)))
        return out

    def time_leaf_hash(self, enable=True):
        _check(lib().p2b_ctx_time_leaf_hash(self.handle, 1 if enable else 0))

    def leaf_hash_time(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().p2b_ctx_leaf_hash_time(self.handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def timer_start(self):
        _check(lib().p2b_timer_start(self.handle))

    def timer_stop_ms(self):
        ms = C.c_float()
        _check(lib().p2b_timer_stop_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    # ---- building blocks ----
    def field_op(self, op, a, b):
        a, b = _np(a), _np(b)
        da, db = DeviceBuffer.from_host(self, a), DeviceBuffer.from_host(self, b)
        do = DeviceBuffer(self, a.size)
        _check(lib().p2b_field_op(self.handle, {"add": 0, "sub": 1, "mul": 2, "mul_add": 3, "add_canonical": 4, "sub_canonical": 5, "add_optimistic": 6, "sub_optimistic": 7,
                                                      "add_optimistic_flag": 8, "sub_optimistic_flag": 9}[op], da.ptr, db.ptr, do.ptr, a.size))
        return do.to_host()

    def poseidon(self, states):
        s = _np(states).reshape(-1, 12)
        d = DeviceBuffer.from_host(self, s)
        _check(lib().p2b_poseidon_permute(self.handle, d.ptr, s.shape[0]))
        return d.to_host().reshape(-1, 12)

    def ifft(self, values):
        """values [P][n] -> coefficients [P][n]  (values.into_par_iter().map(|v| v.ifft()))"""
        v = _np(values)
        P, n = v.shape
        d = DeviceBuffer.from_host(self, v)
        _check(lib().p2b_ifft_batch(self.handle, d.ptr, d.ptr, n.bit_length() - 1, P))
        return d.to_host().reshape(P, n)

    def lde_leaves(self, coeffs, rate_bits):
        """coeffs [P][n] -> leaves [N][P] in the reference's leaf order"""
        c = _np(coeffs)
        P, n = c.shape
        N = n << rate_bits
        d = DeviceBuffer.from_host(self, c)
        out = DeviceBuffer(self, N * P)
        _check(lib().p2b_lde_leaves(self.handle, d.ptr, n.bit_length() - 1, P, rate_bits, out.ptr, P, 0))
        return out.to_host().reshape(N, P)

    def merkle_tree(self, leaves, cap_height):
        """MerkleTree::new(leaves, cap_height) -> (digests [num_digests][4], cap [2^cap_height][4])"""
        lv = _np(leaves)
        N, ll = lv.shape
        ncap = 1 << cap_height
        nd = max(2 * (N - ncap), 0)
        d = DeviceBuffer.from_host(self, lv)
        dd, dc = DeviceBuffer(self, max(nd, 1) * 4), DeviceBuffer(self, max(ncap, 1) * 4)
        _check(lib().p2b_merkle_tree(self.handle, d.ptr, N, ll, ll, 1, cap_height, dd.ptr, dc.ptr))
        return dd.to_host(nd * 4).reshape(nd, 4), dc.to_host(ncap * 4).reshape(ncap, 4)

    def fill_synthetic(self, dbuf, count, seed, first_index=0):
        _check(lib().p2b_fill_synthetic(self.handle, dbuf.ptr, count, seed, first_index))


def compute_quotient_polys(ctx, circuit, wires, zs_partial_products, constants_sigmas, public_inputs_hash, betas, gammas, alphas):
    """compute_quotient_polys (plonk/prover.rs:790-1034) on three committed PolynomialBatch objects.
    Returns (values [num_challenges][lde_size], coeffs [num_challenges][lde_size]) as numpy arrays."""
    nc, size = circuit.num_challenges, circuit.lde_size
    dv, dc = DeviceBuffer(ctx, nc * size), DeviceBuffer(ctx, nc * size)
    arr = lambda x, n: (C.c_uint64 * n)(*[int(v) % ORDER for v in x])
    _check(lib().p2b_quotient_polys(ctx.handle, C.byref(circuit.struct), wires.handle, zs_partial_products.handle, constants_sigmas.handle,
                                    arr(public_inputs_hash, 4), arr(betas, nc), arr(gammas, nc), arr(alphas, nc), dv.ptr, dc.ptr))
    ctx.synchronize()
    return dv.to_host().reshape(nc, size), dc.to_host().reshape(nc, size)


def compute_quotient_polys_rows(ctx, circuit, wires_rows, zs_pp_rows, consts_sigmas_rows, public_inputs_hash, betas, gammas, alphas):
    """Same as compute_quotient_polys, on raw LDE rows ([lde_size][width] uint64, bit-reversed row order) instead of committed
    batches -- the shape the reference kernel reads (plonky2_gpu.cu:684-760: d_ext_wires / d_ext_zs / d_ext_constants_sigmas)."""
    nc, size = circuit.num_challenges, circuit.lde_size
    mats = [np.ascontiguousarray(m, dtype=np.uint64) for m in (wires_rows, zs_pp_rows, consts_sigmas_rows)]
    devs = [DeviceBuffer.from_host(ctx, m.reshape(-1)) for m in mats]
    dv, dc = DeviceBuffer(ctx, nc * size), DeviceBuffer(ctx, nc * size)
    arr = lambda x, n: (C.c_uint64 * n)(*[int(v) % ORDER for v in x])
    _check(lib().p2b_quotient_polys_rows(ctx.handle, C.byref(circuit.struct), devs[0].ptr, mats[0].shape[1], devs[1].ptr, mats[1].shape[1],
                                         devs[2].ptr, mats[2].shape[1], arr(public_inputs_hash, 4), arr(betas, nc), arr(gammas, nc),
                                         arr(alphas, nc), dv.ptr, dc.ptr))
    ctx.synchronize()
    return dv.to_host().reshape(nc, size), dc.to_host().reshape(nc, size)


def partial_products_and_zs(ctx, wires_values, sigma_values, k_is, betas, gammas, quotient_degree_factor):
    """all_wires_permutation_partial_products + the Z-first ordering (plonk/prover.rs:702-786, :112-117).
    wires_values [num_wires][n], sigma_values [num_routed][n]: numpy arrays or (DeviceBuffer, rows, n).
    Returns a DeviceBuffer holding [num_challenges * ceil(num_routed / qdf)][n] and its shape."""
    def dev(x):
        if isinstance(x, tuple):
            return x
        x = _np(x)
        return DeviceBuffer.from_host(ctx, x.reshape(-1)), x.shape[0], x.shape[1]
    (dw, _, n), (ds, nr, n2) = dev(wires_values), dev(sigma_values)
    assert n == n2 and nr == len(k_is)
    nc = len(betas)
    K = -(-nr // quotient_degree_factor)
    out = DeviceBuffer(ctx, max(nc * K, 1) * n)
    arr = lambda x: (C.c_uint64 * max(len(x), 1))(*[int(v) % ORDER for v in x])
    _check(lib().p2b_partial_products_and_zs(ctx.handle, dw.ptr, ds.ptr, n.bit_length() - 1, nr, quotient_degree_factor, nc, arr(k_is),
                                             arr(betas), arr(gammas), out.ptr))
    return out, (nc * K, n)


class Challenger:
    """Host view of the transcript state the FRI calls advance (iop/challenger.rs:15-150).  Only the state lives here:
    permutations run on the device inside p2b_fri_prove_openings."""

    def __init__(self, sponge_state=None, input_buffer=(), output_buffer=()):
        self.struct = ChallengerStruct()
        for k, v in enumerate(sponge_state if sponge_state is not None else [0] * 12):
            self.struct.sponge_state[k] = int(v) % ORDER
        for k, v in enumerate(input_buffer):
            self.struct.input_buffer[k] = int(v) % ORDER
        for k, v in enumerate(output_buffer):
            self.struct.output_buffer[k] = int(v) % ORDER
        self.struct.input_len, self.struct.output_len = len(input_buffer), len(output_buffer)

    @property
    def sponge_state(self):
        return [int(x) for x in self.struct.sponge_state]

    @property
    def input_buffer(self):
        return [int(x) for x in self.struct.input_buffer][:self.struct.input_len]

    @property
    def output_buffer(self):
        return [int(x) for x in self.struct.output_buffer][:self.struct.output_len]


class FriProof:
    """FriProof (fri/proof.rs) as produced by p2b_fri_prove_openings; arrays are numpy uint64."""

    def __init__(self, handle, arity_bits, oracle_leaf_lens, oracle_depths):
        self.handle = handle
        self.arity_bits = list(arity_bits)
        info = FriProofInfo()
        _check(lib().p2b_fri_proof_get_info(handle, C.byref(info)))
        self.info = info
        R, Q, ncap = info.num_reductions, info.num_query_rounds, 1 << info.cap_height
        L = lib()
        get = lambda fn, shape, *a: (lambda arr: (_check(fn(handle, *a, arr.ctypes.data_as(C.c_void_p))), arr)[1])(
            np.zeros(shape, dtype=np.uint64))
        self.commit_phase_merkle_caps = [get(L.p2b_fri_proof_get_cap, (ncap, 4), r) for r in range(R)]
        self.final_poly = get(L.p2b_fri_proof_get_final_poly, (info.final_poly_len, 2))
        self.pow_witness = int(get(L.p2b_fri_proof_get_pow_witness, (1,))[0])
        self.query_indices = [int(x) for x in get(L.p2b_fri_proof_get_query_indices, (max(Q, 1),))[:Q]]
        self.alpha = tuple(int(x) for x in get(L.p2b_fri_proof_get_debug, (2,), 0))
        self.betas = [tuple(int(x) for x in b) for b in get(L.p2b_fri_proof_get_debug, (max(R, 1), 2), 1)[:R]]
        self.pow_response = int(get(L.p2b_fri_proof_get_debug, (1,), 3)[0])
        self.initial = []   # per oracle: (rows [Q][leaf_len], siblings [Q][depth][4])
        for o, (ll, d) in enumerate(zip(oracle_leaf_lens, oracle_depths)):
            rows, sibs = np.zeros((Q, ll), dtype=np.uint64), np.zeros((Q, d, 4), dtype=np.uint64)
            _check(L.p2b_fri_proof_get_initial(handle, o, rows.ctypes.data_as(C.c_void_p), sibs.ctypes.data_as(C.c_void_p)))
            self.initial.append((rows, sibs))
        self.steps = []     # per reduction: (evals [Q][arity][2], siblings [Q][depth][4])
        for r in range(R):
            depth = C.c_uint32()
            _check(L.p2b_fri_proof_get_step(handle, r, None, None, C.byref(depth)))
            ev, sibs = np.zeros((Q, 1 << self.arity_bits[r], 2), dtype=np.uint64), np.zeros((Q, depth.value, 4), dtype=np.uint64)
            _check(L.p2b_fri_proof_get_step(handle, r, ev.ctypes.data_as(C.c_void_p), sibs.ctypes.data_as(C.c_void_p), None))
            self.steps.append((ev, sibs))

    def final_poly_in(self, n):
        """The polynomial that enters FRI (fri/oracle.rs:1084), [n][2] -- parity-test accessor."""
        out = np.zeros((n, 2), dtype=np.uint64)
        _check(lib().p2b_fri_proof_get_debug(self.handle, 2, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if self.handle:
            lib().p2b_fri_proof_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eval_openings(ctx, batch, point):
    """eval_commitment of OpeningSet::new (plonk/proof.rs:313-319): [num_polys][2]."""
    out = np.zeros((batch.num_polys, 2), dtype=np.uint64)
    pt = (C.c_uint64 * 2)(int(point[0]) % ORDER, int(point[1]) % ORDER)
    _check(lib().p2b_eval_openings(ctx.handle, batch.handle, pt, out.ctypes.data_as(C.c_void_p)))
    return out


def fri_prove_openings(ctx, oracles, batches, challenger, degree_bits, rate_bits, cap_height, proof_of_work_bits,
                       num_query_rounds, reduction_arity_bits):
    """PolynomialBatch::prove_openings (fri/oracle.rs:1046-1110).  oracles: PolynomialBatch list; batches: list of
    (point (c0, c1), [(oracle_index, polynomial_index), ...]); challenger: Challenger (advanced in place)."""
    keep = []
    bstructs = (FriBatchInfoStruct * max(len(batches), 1))()
    for k, (point, polys) in enumerate(batches):
        arr = (FriPolyInfo * max(len(polys), 1))()
        for j, (o, p) in enumerate(polys):
            arr[j].oracle_index, arr[j].polynomial_index = o, p
        keep.append(arr)
        bstructs[k].point[0], bstructs[k].point[1] = int(point[0]) % ORDER, int(point[1]) % ORDER
        bstructs[k].polynomials, bstructs[k].num_polynomials = arr, len(polys)
    ab = (C.c_uint32 * max(len(reduction_arity_bits), 1))(*reduction_arity_bits)
    params = FriParamsStruct(degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, len(reduction_arity_bits), ab)
    handles = (C.c_void_p * max(len(oracles), 1))(*[o.handle for o in oracles])
    out = C.c_void_p()
    _check(lib().p2b_fri_prove_openings(ctx.handle, handles, len(oracles), bstructs, len(batches), C.byref(challenger.struct),
                                        C.byref(params), C.byref(out)))
    return FriProof(out, reduction_arity_bits, [o.leaf_len for o in oracles],
                    [o.degree_log + o.rate_bits - o.cap_height for o in oracles])


def _fri_args(batches, degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, reduction_arity_bits):
    keep = []
    bstructs = (FriBatchInfoStruct * max(len(batches), 1))()
    for k, (point, polys) in enumerate(batches):
        arr = (FriPolyInfo * max(len(polys), 1))()
        for j, (o, p) in enumerate(polys):
            arr[j].oracle_index, arr[j].polynomial_index = o, p
        keep.append(arr)
        bstructs[k].point[0], bstructs[k].point[1] = int(point[0]) % ORDER, int(point[1]) % ORDER
        bstructs[k].polynomials, bstructs[k].num_polynomials = arr, len(polys)
    ab = (C.c_uint32 * max(len(reduction_arity_bits), 1))(*reduction_arity_bits)
    keep.append(ab)
    params = FriParamsStruct(degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, len(reduction_arity_bits), ab)
    return bstructs, params, keep


class MerkleTree:
    """View of a committed batch's tree (reference: MerkleTree {leaves, digests, cap}, merkle_tree.rs:41-66)."""

    def __init__(self, batch):
        self._b = batch

    @property
    def cap(self):
        return self._b.cap()

    @property
    def digests(self):
        return self._b.digests()

    def get(self, i):
        return self._b.leaves(i, 1)[0]

    def prove(self, leaf_index):
        return self._b.prove([leaf_index])[0]


class PolynomialBatch:
    """PolynomialBatch (fri/oracle.rs:112-120) resident on the device."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        info = BatchInfo()
        _check(lib().p2b_batch_get_info(handle, C.byref(info)))
        self.degree_log, self.rate_bits, self.cap_height = info.degree_log, info.rate_bits, info.cap_height
        self.salt_size, self.num_polys, self.num_leaves = info.salt_size, info.num_polys, info.num_leaves
        self.leaf_len, self.num_digests = info.leaf_len, info.num_digests
        self.blinding = info.salt_size != 0
        self.merkle_tree = MerkleTree(self)

    @staticmethod
    def _commit(fn, ctx, x, rate_bits, cap_height, salt, blinding):
        if blinding and salt is None:
            raise ValueError("blinding=True needs the salt columns [4][N] (the reference draws them at random)")
        h = C.c_void_p()
        if isinstance(x, DeviceBuffer):
            raise TypeError("pass (DeviceBuffer, P, n) tuples for device-resident input")
        if isinstance(x, tuple):
            dbuf, P, n = x
            ptr, on_host = dbuf.ptr, 0
        else:
            x = _np(x)
            if x.ndim != 2 or x.shape[0] == 0:
                raise P2BError(P2B_ERR_INVALID, "empty batch (no polynomials)")
            P, n = x.shape
            ptr, on_host = x.ctypes.data, 1
        if n == 0 or n & (n - 1):
            raise P2BError(P2B_ERR_INVALID, "polynomial length must be a power of two")
        sp, s_host = None, 1
        if salt is not None:
            if isinstance(salt, DeviceBuffer):
                sp, s_host = salt.ptr, 0
            else:
                salt = _np(salt)
                assert salt.shape == (SALT_SIZE, n << rate_bits)
                sp = salt.ctypes.data
        _check(fn(ctx.handle, ptr, on_host, n.bit_length() - 1, P, rate_bits, cap_height, sp, s_host, C.byref(h)))
        return PolynomialBatch(ctx, h.value)

    @classmethod
    def from_values(cls, ctx, values, rate_bits, cap_height, blinding=False, salt=None, coeffs_out=None):
        """fri/oracle.rs:709-731.  values: [P][n] numpy array (host) or (DeviceBuffer, P, n).
        coeffs_out: optional pinned host array [P][n] that receives the coefficients while the tree is built."""
        if coeffs_out is not None:
            ptr = coeffs_out.ctypes.data
            fn = lambda *a: lib().p2b_commit_from_values_ex(*a[:-1], ptr, a[-1])
            return cls._commit(fn, ctx, values, rate_bits, cap_height, salt, blinding)
        return cls._commit(lib().p2b_commit_from_values, ctx, values, rate_bits, cap_height, salt, blinding)

    @classmethod
    def from_coeffs(cls, ctx, coeffs, rate_bits, cap_height, blinding=False, salt=None):
        """fri/oracle.rs:911-977."""
        return cls._commit(lib().p2b_commit_from_coeffs, ctx, coeffs, rate_bits, cap_height, salt, blinding)

    def close(self):
        if self.handle:
            lib().p2b_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard_info(self):
        a, b, t, f, n = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(lib().p2b_batch_shard_info(self.handle, C.byref(a), C.byref(b), C.byref(t), C.byref(f), C.byref(n)))
        return {"first_leaf": a.value, "local_leaves": b.value, "top_layer": t.value, "top_node_first": f.value, "top_node_count": n.value}

    def device_ptrs(self):
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().p2b_batch_device_ptrs(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"coeffs": a.value, "leaves": b.value, "digests": c.value, "cap": d.value}

    def polynomials(self, out=None):
        out = np.empty((self.num_polys, 1 << self.degree_log), dtype=np.uint64) if out is None else out
        _check(lib().p2b_batch_get_coeffs(self.handle, out.ctypes.data))
        return out

    def cap(self, out=None):
        out = np.empty((1 << self.cap_height, 4), dtype=np.uint64) if out is None else out
        _check(lib().p2b_batch_get_cap(self.handle, out.ctypes.data))
        return out

    def digests(self):
        out = np.empty((self.num_digests, 4), dtype=np.uint64)
        if self.num_digests:
            _check(lib().p2b_batch_get_digests(self.handle, out.ctypes.data))
        return out

    def leaves(self, first=None, count=None):
        if first is None or count is None:
            si = self.shard_info()
            first = si["first_leaf"] if first is None else first
            count = si["first_leaf"] + si["local_leaves"] - first if count is None else count
        out = np.empty((count, self.leaf_len), dtype=np.uint64)
        _check(lib().p2b_batch_get_leaves(self.handle, first, count, out.ctypes.data))
        return out

    def get_lde_values(self, index, step=1):
        """fri/oracle.rs:1007-1018"""
        out = np.empty(self.num_polys, dtype=np.uint64)
        _check(lib().p2b_batch_get_lde_values(self.handle, index, step, out.ctypes.data))
        return out

    def prove(self, leaf_indices):
        idx = _np(leaf_indices)
        layers = self.degree_log + self.rate_bits - self.cap_height
        out = np.empty((idx.size, layers, 4), dtype=np.uint64)
        _check(lib().p2b_batch_prove(self.handle, idx.ctypes.data, idx.size, out.ctypes.data if layers else None))
        return out

    def open_rows(self, leaf_indices, with_proofs=True):
        idx = _np(leaf_indices)
        layers = self.degree_log + self.rate_bits - self.cap_height
        rows = np.empty((idx.size, self.leaf_len), dtype=np.uint64)
        sib = np.empty((idx.size, layers, 4), dtype=np.uint64) if with_proofs else None
        _check(lib().p2b_batch_open_rows(self.handle, idx.ctypes.data, idx.size, rows.ctypes.data,
                                         sib.ctypes.data if (with_proofs and layers) else None))
        return rows, sib


class MultiGpu:
    """p2b_mgpu: ONE process driving several devices (the shape of the reference's caller, fri/oracle.rs:279-545)."""

    def __init__(self, devices=None, count=None):
        if devices is None:
            devices = list(range(count if count is not None else 1))
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().p2b_mgpu_create(arr, len(devices), C.byref(h)))
        self.handle, self.devices = h, list(devices)

    def close(self):
        if self.handle:
            lib().p2b_mgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def peer_access(self):
        return bool(lib().p2b_mgpu_peer_access(self.handle))

    def synchronize(self):
        _check(lib().p2b_mgpu_synchronize(self.handle))

    def timer_start(self):
        _check(lib().p2b_mgpu_timer_start(self.handle))

    def timer_stop_ms(self):
        ms = C.c_float()
        _check(lib().p2b_mgpu_timer_stop_ms(self.handle, C.byref(ms)))
        return ms.value

    def commit_from_values(self, values, rate_bits, cap_height, coeffs_out=None):
        """values: host array [P][n] (numpy; pinned memory via PinnedBuffer overlaps the upload)."""
        v = values if (isinstance(values, np.ndarray) and values.dtype == np.uint64 and values.flags.c_contiguous) else _np(values)
        P, n = v.shape
        h = C.c_void_p()
        _check(lib().p2b_mgpu_commit_from_values(self.handle, v.ctypes.data, n.bit_length() - 1, P, rate_bits, cap_height,
                                                 coeffs_out.ctypes.data if coeffs_out is not None else None, C.byref(h)))
        b = MultiGpuBatch(self, h)
        b._keep = (v, coeffs_out)
        return b

    def resident_cols(self, index, degree_log, num_polys):
        ptr, rounds = C.c_void_p(), C.c_uint64()
        _check(lib().p2b_mgpu_resident_cols(self.handle, index, degree_log, num_polys, C.byref(ptr), C.byref(rounds)))
        return ptr.value, rounds.value

    def _ctx(self, index):
        return lib().p2b_mgpu_ctx(self.handle, index)

    def context(self, index):
        """Device `index`'s context as a Context object (owned by the group: closing it is a no-op)."""
        return Context(_borrowed=self._ctx(index))

    def commit_from_device_values(self, src_index, ptr, degree_log, num_polys, rate_bits, cap_height):
        """ptr: device pointer on device `src_index` to the value matrix [num_polys][n]."""
        h = C.c_void_p()
        _check(lib().p2b_mgpu_commit_from_device_values(self.handle, src_index, C.c_void_p(ptr), degree_log, num_polys, rate_bits, cap_height, C.byref(h)))
        return MultiGpuBatch(self, h)

    def commit_from_device_coeffs(self, ptrs, degree_log, num_polys, rate_bits, cap_height):
        """ptrs[d]: device d's pointer to the full coefficient matrix [num_polys][n]."""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        h = C.c_void_p()
        _check(lib().p2b_mgpu_commit_from_device_coeffs(self.handle, arr, degree_log, num_polys, rate_bits, cap_height, C.byref(h)))
        return MultiGpuBatch(self, h)

    def quotient_polys(self, circuit, wires, zs_pp, consts_sigmas, pih, betas, gammas, alphas):
        """compute_quotient_polys over sharded batches; returns per-device pointers to [num_challenges][lde_size] coefficients
        (allocated with p2b_malloc on each device; free with free_device_ptrs)."""
        L = lib()
        n_words = circuit.num_challenges * circuit.lde_size
        ptrs = []
        for d in range(len(self.devices)):
            p = C.c_void_p()
            _check(L.p2b_malloc(self._ctx(d), n_words * 8, C.byref(p)))
            ptrs.append(p.value)
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        a = lambda x: (C.c_uint64 * len(x))(*[int(v) % ORDER for v in x])   # noqa: E731
        _check(L.p2b_mgpu_quotient_polys(self.handle, C.byref(circuit.struct), wires.handle, zs_pp.handle, consts_sigmas.handle, a(pih), a(betas),
                                         a(gammas), a(alphas), arr))
        return ptrs

    def read_device(self, index, ptr, count):
        out = np.empty(count, dtype=np.uint64)
        _check(lib().p2b_memcpy_d2h(self._ctx(index), out.ctypes.data, C.c_void_p(ptr), count * 8))
        return out

    def free_device_ptrs(self, ptrs):
        self.synchronize()
        for d, p in enumerate(ptrs):
            _check(lib().p2b_free(self._ctx(d), C.c_void_p(p)))

    def eval_openings(self, batch, point):
        out = np.empty((batch.info.num_polys, 2), dtype=np.uint64)
        pt = (C.c_uint64 * 2)(int(point[0]) % ORDER, int(point[1]) % ORDER)
        _check(lib().p2b_mgpu_eval_openings(self.handle, batch.handle, pt, out.ctypes.data))
        return out

    def fri_prove_openings(self, oracles, batches, challenger, degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds,
                           reduction_arity_bits):
        bstructs, params, keep = _fri_args(batches, degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, reduction_arity_bits)
        handles = (C.c_void_p * max(len(oracles), 1))(*[o.handle for o in oracles])
        out = C.c_void_p()
        _check(lib().p2b_mgpu_fri_prove_openings(self.handle, handles, len(oracles), bstructs, len(batches), C.byref(challenger.struct),
                                                 C.byref(params), C.byref(out)))
        return FriProof(out, reduction_arity_bits, [o.leaf_len for o in oracles], [o.layers for o in oracles])

    def commit_resident(self, degree_log, num_polys, rate_bits, cap_height):
        h = C.c_void_p()
        _check(lib().p2b_mgpu_commit_resident(self.handle, degree_log, num_polys, rate_bits, cap_height, None, C.byref(h)))
        return MultiGpuBatch(self, h)


class MultiGpuBatch:
    def __init__(self, mg, handle):
        self.mg, self.handle = mg, handle
        info = BatchInfo()
        _check(lib().p2b_mgpu_batch_get_info(handle, C.byref(info)))
        self.info = info
        self.leaf_len, self.num_leaves, self.cap_height = info.leaf_len, info.num_leaves, info.cap_height
        self.layers = info.degree_log + info.rate_bits - info.cap_height

    def close(self):
        if self.handle:
            lib().p2b_mgpu_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cap(self):
        out = np.empty((1 << self.cap_height, 4), dtype=np.uint64)
        _check(lib().p2b_mgpu_batch_get_cap(self.handle, out.ctypes.data))
        return out

    def leaves(self, first=0, count=None):
        count = self.num_leaves - first if count is None else count
        out = np.empty((count, self.leaf_len), dtype=np.uint64)
        _check(lib().p2b_mgpu_batch_get_leaves(self.handle, first, count, out.ctypes.data))
        return out

    def open_rows(self, leaf_indices):
        idx = _np(leaf_indices)
        rows = np.empty((idx.size, self.leaf_len), dtype=np.uint64)
        sib = np.empty((idx.size, self.layers, 4), dtype=np.uint64)
        _check(lib().p2b_mgpu_batch_open_rows(self.handle, idx.ctypes.data, idx.size, rows.ctypes.data, sib.ctypes.data if self.layers else None))
        return rows, sib
