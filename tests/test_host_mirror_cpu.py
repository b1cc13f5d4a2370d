"""CPU: the C++ host mirror (plonky2-gpu_b200/host/polynomial_batch.hpp) compiles against the C ABI and links."""
import os
import subprocess
import tempfile

import plonky2_gpu_b200 as p2b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "plonky2-gpu_b200/host/polynomial_batch.hpp"
#include <cstdio>
int main() {
  using namespace plonky2_b200;
  try {
    Context ctx(0);
    std::vector<std::vector<F>> vals(3, std::vector<F>(8, 1));
    PolynomialBatch b = PolynomialBatch::from_values(ctx, vals, 1, false, 0);
    MerkleCap cap = b.merkle_tree.cap();
    // the FRI-facing mirror: openings and prove_openings over the same batch
    Ext zeta{{5, 7}};
    std::vector<Ext> op = eval_openings(ctx, b, zeta);
    Challenger ch;
    FriBatchInfo fb{zeta, {{0, 0}, {0, 1}, {0, 2}}};
    FriProof pr = prove_openings(ctx, {&b}, {fb}, ch, 3, 1, 0, 2, 3, {1});
    if (pr.query_round_proofs.size() != 3 || pr.final_poly.size() != 4 || op.size() != 3) return 4;
    std::printf("cap %llu\n", (unsigned long long)cap.hashes[0].elements[0]);
  } catch (const std::exception& e) { std::printf("ERR %s\n", e.what()); return 3; }
  return 0;
}
'''


def test_host_mirror_compiles_and_links():
    so = p2b.build()
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(src, "w").write(SRC)
        cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I" + ROOT, src, "-o", exe, "-L" + os.path.dirname(so), "-lplonky2_b200",
               "-Wl,-rpath," + os.path.dirname(so), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        # without a GPU the program must fail loudly through the library's error path, not fall back
        run = subprocess.run([exe], capture_output=True, text=True)
        import torch
        if not torch.cuda.is_available():
            assert run.returncode == 3 and "ERR plonky2_b200" in run.stdout, run.stdout + run.stderr
        else:
            assert run.returncode == 0 and run.stdout.startswith("cap "), run.stdout + run.stderr
