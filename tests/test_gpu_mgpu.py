"""GPU: the single-process multi-device commit (p2b_mgpu_*, include/plonky2_b200.h) -- the entry a one-process prover
(plonky2/src/fri/oracle.rs:279-545) binds.  With one visible GPU the same flow runs with n_dev = 1 (rounds, peer push to
itself, progressive absorb); with >= 2 GPUs it runs over 2 (and 4, 8 when present)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _counts():
    return [g for g in (1, 2, 4, 8)]


@pytest.mark.parametrize("ndev", _counts())
@pytest.mark.parametrize("n_log,P,cap_height", [(10, 135, 4), (9, 21, 0), (12, 9, 2)])
def test_mgpu_commit_matches_oracle(ndev, n_log, P, cap_height):
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    p2b.build()
    rng = np.random.default_rng(100 + n_log)
    values = rng.integers(0, oracle.ORDER, size=(P, 1 << n_log), dtype=np.uint64)
    want = oracle.batch_from_values(values, 3, cap_height)
    mg = p2b.MultiGpu(count=ndev)
    coeffs = np.empty_like(values)
    for _ in range(2):
        b = mg.commit_from_values(values, 3, cap_height, coeffs_out=coeffs)
        assert np.array_equal(b.cap(), want.cap)
        assert np.array_equal(coeffs, want.coeffs)
        assert np.array_equal(b.leaves(), want.leaves)
        idx = [int(x) for x in rng.integers(0, want.leaves.shape[0], size=24)]
        rows, sibs = b.open_rows(idx)
        for k, x in enumerate(idx):
            assert np.array_equal(rows[k], want.leaves[x])
            assert oracle.merkle_verify(rows[k], x, want.cap, sibs[k])
        b.close()
    mg.close()


def test_mgpu_rejects_bad_shapes():
    p2b.build()
    with pytest.raises(p2b.P2BError):
        p2b.MultiGpu(devices=[0, 0])
    mg = p2b.MultiGpu(count=1)
    with pytest.raises(p2b.P2BError):
        mg.commit_from_values(np.zeros((3, 16), dtype=np.uint64), 3, 0)   # <= 4 polynomials: hash_or_noop copies
    mg.close()


def test_c_program_drives_the_devices_from_one_process(tmp_path):
    """tests/c/mgpu_commit_test.c: plain C against include/plonky2_b200.h, no Python in the data path."""
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    p2b.build()
    exe = str(tmp_path / "mgpu_commit_test")
    libdir = os.path.join(ROOT, "plonky2-gpu_b200")
    subprocess.run([cc, "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "mgpu_commit_test.c"), "-o", exe,
                    "-L", libdir, "-lplonky2_b200", "-Wl,-rpath," + libdir], check=True)
    ndev = 2 if _ngpu() >= 2 else 1
    r = subprocess.run([exe, str(ndev), "12", "135"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK")
