"""Build libplonky2_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

    python plonky2-gpu_b200/build.py [--force] [--verbose]

The library is one translation unit (csrc/plonky2_b200.cu); the .so lands next to this file so it travels
to the GPU box with the repo snapshot.  Mirrors what the reference's cuda/build.rs:19-45 does with the `cc`
crate (one nvcc invocation producing the library the Rust crate links).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libplonky2_b200.so")
SOURCES = ["plonky2_b200.cu"]


def deps():
    """Every file the translation unit can include: all of csrc/ plus the public header (a stale .so must never pass tests)."""
    import glob
    files = []
    for pat in ("*.cu", "*.cuh", "*.h", "*.hpp"):
        files += glob.glob(os.path.join(CSRC, pat))
    files.append(os.path.join(HERE, "..", "include", "plonky2_b200.h"))
    files.append(os.path.abspath(__file__))
    return files


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    for p in deps():
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (tools/variants.py) next to the product library."""
    so = out or SO
    if out is None and not force and not needs_build():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-O3",
           "-o", so] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    # the host compiler of this image's CC/CXX env may lack its specs; use the system one
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libplonky2_b200.so")
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
