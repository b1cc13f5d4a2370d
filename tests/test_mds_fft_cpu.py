"""CPU: the convolution form of the Poseidon MDS layer (plonky2-gpu_b200/csrc/mds_fft.cuh) equals the direct circulant form
(poseidon.rs:172-260) for int64, wrapping uint32 and double arithmetic, on random inputs and on every vertex of the input
box, and every intermediate of the FP64 evaluation stays below 2^53 (exactness of the device path)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mds_fft_matches_direct_form(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "mds_fft_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "mds_fft_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok")
