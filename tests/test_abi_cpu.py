"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/plonky2_b200.h declares.
No compute calls (there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import pytest

import plonky2_gpu_b200 as p2b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return p2b.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "plonky2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_declares_the_reference_ffi_symbols():
    syms = declared_symbols()
    # cuda/src/lib.rs:52-145
    for ref in ["init", "ifft", "build_merkle_tree", "merkle_tree_from_values", "merkle_tree_from_coeffs", "compute_quotient_polys"]:
        assert ref in syms, ref
    assert len([s for s in syms if s.startswith("p2b_")]) >= 30


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_error_string(built):
    L = p2b.lib()
    assert b"sm_100a" in L.p2b_version()
    assert L.p2b_last_error() is not None


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(p2b.P2BError):
        p2b.Context()


def test_product_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under plonky2-gpu_b200/ or include/ may reference it
    bad = []
    for base in ("plonky2-gpu_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"p2oracle|import oracle|from oracle|oracle/", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_header_is_plain_c99():
    """The drop-in boundary is a C ABI: the header must compile as C (no C++-isms, no torch / CUDA types in signatures)."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "h.c")
        open(src, "w").write('#include "include/plonky2_b200.h"\nint main(void) { p2b_challenger c; (void)c; return 0; }\n')
        r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I" + ROOT, src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(os.path.join(ROOT, "include", "plonky2_b200.h")).read()
    assert "torch" not in text and "cuda_runtime" not in text and "cudaStream_t" not in text.split("*/")[-1]
