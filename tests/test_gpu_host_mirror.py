"""GPU: the C++ host mirror (plonky2-gpu_b200/host/polynomial_batch.hpp -- PolynomialBatch / MerkleTree / prove_openings with
the reference's names, fri/oracle.rs:709-1110) EXECUTES on the device and agrees with the CPU oracle: cap, an opened row with
its Merkle proof, get_lde_values."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import plonky2_gpu_b200 as p2b

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "plonky2-gpu_b200/host/polynomial_batch.hpp"
#include <cstdio>
int main() {
  using namespace plonky2_b200;
  try {
    Context ctx(0);
    const int P = 21, n = 256;
    std::vector<std::vector<F>> vals(P, std::vector<F>(n));
    uint64_t s = 12345;
    for (int c = 0; c < P; c++) for (int i = 0; i < n; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; vals[c][i] = (s >> 1) % 0xFFFFFFFF00000001ull; }
    PolynomialBatch b = PolynomialBatch::from_values(ctx, vals, 3, false, 2);
    MerkleCap cap = b.merkle_tree.cap();
    for (auto& h : cap.hashes) std::printf("cap %llu %llu %llu %llu\n", (unsigned long long)h.elements[0], (unsigned long long)h.elements[1], (unsigned long long)h.elements[2], (unsigned long long)h.elements[3]);
    std::vector<F> row = b.merkle_tree.get(777);
    std::printf("row");
    for (F v : row) std::printf(" %llu", (unsigned long long)v);
    std::printf("\n");
    MerkleProof pr = b.merkle_tree.prove(777);
    for (auto& h : pr.siblings) std::printf("sib %llu %llu %llu %llu\n", (unsigned long long)h.elements[0], (unsigned long long)h.elements[1], (unsigned long long)h.elements[2], (unsigned long long)h.elements[3]);
    std::vector<F> lv = b.get_lde_values(5, 8);
    std::printf("lde");
    for (F v : lv) std::printf(" %llu", (unsigned long long)v);
    std::printf("\n");
  } catch (const std::exception& e) { std::printf("ERR %s\n", e.what()); return 3; }
  return 0;
}
'''


def test_host_mirror_runs_on_the_device_and_matches_the_oracle(tmp_path):
    so = p2b.build()
    src, exe = str(tmp_path / "t.cpp"), str(tmp_path / "t")
    open(src, "w").write(SRC)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I" + ROOT, src, "-o", exe, "-L" + os.path.dirname(so), "-lplonky2_b200",
           "-Wl,-rpath," + os.path.dirname(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    P, n = 21, 256
    vals = np.empty((P, n), dtype=np.uint64)
    s = 12345
    for c in range(P):
        for i in range(n):
            s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            vals[c, i] = (s >> 1) % oracle.ORDER
    want = oracle.batch_from_values(vals, 3, 2)
    lines = run.stdout.splitlines()
    cap = np.array([[int(x) for x in l.split()[1:]] for l in lines if l.startswith("cap ")], dtype=np.uint64)
    assert np.array_equal(cap, want.cap)
    row = np.array([int(x) for x in next(l for l in lines if l.startswith("row")).split()[1:]], dtype=np.uint64)
    assert np.array_equal(row, want.leaves[777])
    sibs = np.array([[int(x) for x in l.split()[1:]] for l in lines if l.startswith("sib ")], dtype=np.uint64)
    assert oracle.merkle_verify(row, 777, want.cap, sibs)
    lde = np.array([int(x) for x in next(l for l in lines if l.startswith("lde")).split()[1:]], dtype=np.uint64)
    assert np.array_equal(lde, want.get_lde_values(5, 8))
