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


def _gpu_counts():
    return [g for g in (1, 2, 4, 8)]


@pytest.mark.parametrize("ndev", _gpu_counts())
def test_mgpu_prove_stages_match_single_device(ndev):
    """Quotient polynomials, quotient-chunk commit, openings and the FRI opening proof over SHARDED batches (one process, ndev
    devices) equal the single-device results bit for bit: plonk/prover.rs:884-1034, fri/oracle.rs:1046-1110."""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    from tests import quotient_fixtures as F
    p2b.build()
    gates, groups, sel = F.recursion_gate_set()
    inst = F.build_instance(gates, groups, sel, 6, 135, 80, seed=33)
    circ = inst.circ
    ctx = p2b.Context(0)
    mg = p2b.MultiGpu(count=ndev)
    rb, ch = circ.rate_bits, 2
    mats = [inst.consts_sigmas, inst.wires, inst.zs_pp]
    single = [p2b.PolynomialBatch.from_values(ctx, m, rb, ch) for m in mats]
    multi = [mg.commit_from_values(np.ascontiguousarray(m, dtype=np.uint64), rb, ch) for m in mats]
    for a, b in zip(single, multi):
        assert np.array_equal(a.cap(), b.cap())
    pc = p2b.Circuit([(g.type_id, g.params) for g in circ.gates], circ.selector_indices, circ.groups, circ.num_wires, circ.num_routed_wires,
                     circ.num_constants, circ.k_is, circ.degree_bits, circ.rate_bits, circ.num_challenges, circ.quotient_degree_factor)
    import ctypes as C
    size, nc = pc.lde_size, pc.num_challenges
    dv, dc = p2b.DeviceBuffer(ctx, nc * size), p2b.DeviceBuffer(ctx, nc * size)
    arr = lambda x: (C.c_uint64 * len(x))(*[int(v) for v in x])   # noqa: E731
    p2b._check(p2b.lib().p2b_quotient_polys(ctx.handle, C.byref(pc.struct), single[1].handle, single[2].handle, single[0].handle, arr(inst.pih),
                                            arr(inst.betas), arr(inst.gammas), arr(inst.alphas), dv.ptr, dc.ptr))
    want_coeffs = dc.to_host(nc * size)
    ptrs = mg.quotient_polys(pc, multi[1], multi[2], multi[0], inst.pih, inst.betas, inst.gammas, inst.alphas)
    for d in range(ndev):
        assert np.array_equal(mg.read_device(d, ptrs[d], nc * size), want_coeffs), d
    # quotient chunks: [nc][8 n] coefficients are [nc * 8][n] chunk-major (prover.rs:151-166)
    n_log = circ.degree_bits
    qdf = 8
    bq_s = p2b.PolynomialBatch.from_coeffs(ctx, (dc, nc * qdf, 1 << n_log), rb, ch)
    bq_m = mg.commit_from_device_coeffs(ptrs, n_log, nc * qdf, rb, ch)
    assert np.array_equal(bq_s.cap(), bq_m.cap())
    # openings + FRI proof over the four oracles
    zeta = (0x1234567, 0x7654321)
    for a, b in zip(single + [bq_s], multi + [bq_m]):
        assert np.array_equal(p2b.eval_openings(ctx, a, zeta), mg.eval_openings(b, zeta))
    polys = [m.shape[0] for m in mats] + [nc * qdf]
    all_polys = [(o, p) for o, k in enumerate(polys) for p in range(k)]
    batches = [(zeta, all_polys), ((5, 9), [(2, 0), (2, 1)])]
    args = (n_log, rb, ch, 6, 5, [2, 1])
    c1, c2 = p2b.Challenger(list(range(12)), [3, 4]), p2b.Challenger(list(range(12)), [3, 4])
    ps = p2b.fri_prove_openings(ctx, single + [bq_s], batches, c1, *args)
    pm = mg.fri_prove_openings(multi + [bq_m], batches, c2, *args)
    assert ps.pow_witness == pm.pow_witness and ps.query_indices == pm.query_indices
    assert np.array_equal(ps.final_poly, pm.final_poly)
    assert all(np.array_equal(a, b) for a, b in zip(ps.commit_phase_merkle_caps, pm.commit_phase_merkle_caps))
    for (r1, s1), (r2, s2) in zip(ps.initial, pm.initial):
        assert np.array_equal(r1, r2) and np.array_equal(s1, s2)
    for (e1, s1), (e2, s2) in zip(ps.steps, pm.steps):
        assert np.array_equal(e1, e2) and np.array_equal(s1, s2)
    assert list(c1.struct.sponge_state) == list(c2.struct.sponge_state)
    ps.close()
    pm.close()
    mg.free_device_ptrs(ptrs)
    for b in single + [bq_s] + multi + [bq_m]:
        b.close()
    mg.close()
    ctx.close()
