"""ctypes/numpy front-end of the CPU oracle (oracle/p2oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package; the product (plonky2-gpu_b200/) never does.  See oracle/p2oracle.h for the parity-pinning
statement and the reference file:line each function follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libp2oracle.so")
ORDER = 0xFFFFFFFF00000001

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("p2oracle.c", "p2oracle.h", "poseidon_tables.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libp2oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, sz, un = C.c_uint64, C.c_size_t, C.c_uint
        for name, res, args in [
            ("p2o_add", u64, [u64, u64]), ("p2o_sub", u64, [u64, u64]), ("p2o_mul", u64, [u64, u64]),
            ("p2o_canon", u64, [u64]), ("p2o_exp", u64, [u64, u64]), ("p2o_inverse", u64, [u64]),
            ("p2o_inverse_2exp", u64, [un]), ("p2o_primitive_root_of_unity", u64, [un]),
            ("p2o_reverse_bits", u64, [u64, un]), ("p2o_reverse_index_bits_in_place", None, [_u64p, sz]),
            ("p2o_root_table_len", sz, [un]), ("p2o_fft_root_table_concat", None, [un, _u64p]),
            ("p2o_fft", None, [_u64p, un, un]), ("p2o_ifft", None, [_u64p, un]),
            ("p2o_coset_fft", None, [_u64p, un, u64, un]), ("p2o_coset_ifft", None, [_u64p, un, u64]),
            ("p2o_lde_coset_fft", None, [_u64p, un, un, _u64p]),
            ("p2o_naive_coset_eval", None, [_u64p, sz, un, u64, _u64p]),
            ("p2o_poseidon", None, [_u64p]), ("p2o_poseidon_naive", None, [_u64p]),
            ("p2o_hash_no_pad", None, [_u64p, sz, _u64p]), ("p2o_hash_or_noop", None, [_u64p, sz, _u64p]),
            ("p2o_two_to_one", None, [_u64p, _u64p, _u64p]),
            ("p2o_merkle_tree", C.c_int, [_u64p, sz, sz, un, _u64p, _u64p]),
            ("p2o_merkle_prove", None, [_u64p, sz, un, sz, _u64p]),
            ("p2o_merkle_verify", C.c_int, [_u64p, sz, sz, _u64p, un, _u64p, sz]),
            ("p2o_batch_from_values", C.c_int, [_u64p, un, sz, un, un, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
            ("p2o_batch_from_coeffs", C.c_int, [_u64p, un, sz, un, un, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
            ("p2o_set_threads", None, [C.c_int]), ("p2o_get_threads", C.c_int, []),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _a(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- field ----
def add(a, b): return int(lib().p2o_add(a, b))
def sub(a, b): return int(lib().p2o_sub(a, b))
def mul(a, b): return int(lib().p2o_mul(a, b))
def exp(a, e): return int(lib().p2o_exp(a, e))
def inverse(a): return int(lib().p2o_inverse(a))
def inverse_2exp(e): return int(lib().p2o_inverse_2exp(e))
def primitive_root_of_unity(k): return int(lib().p2o_primitive_root_of_unity(k))
def reverse_bits(n, bits): return int(lib().p2o_reverse_bits(n, bits))
def set_threads(n): lib().p2o_set_threads(int(n))
def get_threads(): return int(lib().p2o_get_threads())


def reverse_index_bits(v):
    v = _a(v).copy()
    lib().p2o_reverse_index_bits_in_place(v, v.size)
    return v


def fft_root_table_concat(n_log):
    out = np.empty(lib().p2o_root_table_len(n_log), dtype=np.uint64)
    lib().p2o_fft_root_table_concat(n_log, out)
    return out


def fft(coeffs, zero_factor=0):
    v = _a(coeffs).copy()
    lib().p2o_fft(v, int(v.size).bit_length() - 1, zero_factor)
    return v


def ifft(values):
    v = _a(values).copy()
    lib().p2o_ifft(v, int(v.size).bit_length() - 1)
    return v


def coset_fft(coeffs, shift=7, zero_factor=0):
    v = _a(coeffs).copy()
    lib().p2o_coset_fft(v, int(v.size).bit_length() - 1, shift, zero_factor)
    return v


def coset_ifft(values, shift=7):
    v = _a(values).copy()
    lib().p2o_coset_ifft(v, int(v.size).bit_length() - 1, shift)
    return v


def lde_coset_fft(coeffs, rate_bits):
    c = _a(coeffs)
    out = np.empty(c.size << rate_bits, dtype=np.uint64)
    lib().p2o_lde_coset_fft(c, int(c.size).bit_length() - 1, rate_bits, out)
    return out


def naive_coset_eval(coeffs, n_log, shift=1):
    c = _a(coeffs)
    out = np.empty(1 << n_log, dtype=np.uint64)
    lib().p2o_naive_coset_eval(c, c.size, n_log, shift, out)
    return out


# ---- Poseidon ----
def poseidon(state, naive=False):
    s = _a(state).copy()
    assert s.size == 12
    (lib().p2o_poseidon_naive if naive else lib().p2o_poseidon)(s)
    return s


def hash_no_pad(x):
    x = _a(x)
    out = np.empty(4, dtype=np.uint64)
    lib().p2o_hash_no_pad(x if x.size else np.zeros(1, np.uint64), x.size, out)
    return out


def hash_or_noop(x):
    x = _a(x)
    out = np.empty(4, dtype=np.uint64)
    lib().p2o_hash_or_noop(x if x.size else np.zeros(1, np.uint64), x.size, out)
    return out


def two_to_one(l, r):
    out = np.empty(4, dtype=np.uint64)
    lib().p2o_two_to_one(_a(l), _a(r), out)
    return out


# ---- Merkle ----
def merkle_tree(leaves, cap_height):
    """leaves: [num_leaves, leaf_len] -> (digests [num_digests,4], cap [2^cap_height,4])"""
    lv = _a(leaves)
    n, ll = lv.shape
    ncap = 1 << cap_height
    nd = 2 * (n - ncap)
    digests = np.empty((max(nd, 0), 4), dtype=np.uint64)
    cap = np.empty((ncap, 4), dtype=np.uint64)
    dg = digests if nd > 0 else np.empty((1, 4), dtype=np.uint64)
    rc = lib().p2o_merkle_tree(lv.reshape(-1) if lv.size else np.zeros(1, np.uint64), n, ll, cap_height, dg.reshape(-1), cap.reshape(-1))
    if rc != 0:
        raise ValueError("cap_height=%d should be at most log2(leaves.len())" % cap_height)
    return digests, cap


def merkle_prove(digests, num_leaves, cap_height, leaf_index):
    layers = (num_leaves.bit_length() - 1) - cap_height
    sib = np.empty((max(layers, 0), 4), dtype=np.uint64)
    if layers > 0:
        lib().p2o_merkle_prove(_a(digests).reshape(-1), num_leaves, cap_height, leaf_index, sib.reshape(-1))
    return sib


def merkle_verify(leaf, leaf_index, cap, siblings):
    leaf, cap, sib = _a(leaf), _a(cap), _a(siblings)
    cap_height = int(cap.shape[0]).bit_length() - 1
    s = sib.reshape(-1) if sib.size else np.zeros(1, np.uint64)
    return bool(lib().p2o_merkle_verify(leaf, leaf.size, leaf_index, cap.reshape(-1), cap_height, s, sib.shape[0]))


# ---- PolynomialBatch ----
class Batch:
    """Result of PolynomialBatch::from_values/from_coeffs on the CPU (fri/oracle.rs:112-120)."""

    def __init__(self, coeffs, leaves, digests, cap, n_log, rate_bits, salt_size):
        self.coeffs, self.leaves, self.digests, self.cap = coeffs, leaves, digests, cap
        self.cap_height = int(cap.shape[0]).bit_length() - 1
        self.degree_log, self.rate_bits, self.salt_size = n_log, rate_bits, salt_size

    def get_lde_values(self, index, step=1):
        """fri/oracle.rs:1007-1018"""
        i = reverse_bits(index * step, self.degree_log + self.rate_bits)
        row = self.leaves[i]
        return row[: row.size - self.salt_size]


def _batch(values_or_coeffs, rate_bits, cap_height, salt, from_values, want_leaves=True, want_digests=True):
    x = _a(values_or_coeffs)
    P, n = x.shape
    n_log = n.bit_length() - 1
    N = n << rate_bits
    salt_size = 0
    if salt is not None:
        salt = _a(salt)
        assert salt.shape == (4, N)
        salt_size = 4
    ncap = 1 << cap_height
    nd = 2 * (N - ncap)
    leaves = np.empty((N, P + salt_size), dtype=np.uint64) if want_leaves else None
    digests = np.empty((max(nd, 1), 4), dtype=np.uint64) if want_digests else None
    cap = np.empty((ncap, 4), dtype=np.uint64)
    if from_values:
        coeffs = np.empty((P, n), dtype=np.uint64)
        rc = lib().p2o_batch_from_values(x.reshape(-1), n_log, P, rate_bits, cap_height, _ptr(salt), _ptr(coeffs), _ptr(leaves), _ptr(digests), _ptr(cap))
    else:
        coeffs = x
        rc = lib().p2o_batch_from_coeffs(x.reshape(-1), n_log, P, rate_bits, cap_height, _ptr(salt), _ptr(leaves), _ptr(digests), _ptr(cap))
    if rc != 0:
        raise ValueError("p2oracle batch failed rc=%d" % rc)
    if digests is not None:
        digests = digests[: max(nd, 0)]
    return Batch(coeffs, leaves, digests, cap, n_log, rate_bits, salt_size)


def batch_from_values(values, rate_bits, cap_height, salt=None, **kw):
    return _batch(values, rate_bits, cap_height, salt, True, **kw)


def batch_from_coeffs(coeffs, rate_bits, cap_height, salt=None, **kw):
    return _batch(coeffs, rate_bits, cap_height, salt, False, **kw)
