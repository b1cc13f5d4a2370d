"""CPU: the cross-check tooling for a machine with the Rust toolchain (tools/rust_crosscheck/) is self-consistent: the
exporter writes every case and the comparer accepts the oracle's own outputs in the dumper's file layout."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_export_and_self_compare(tmp_path):
    d = str(tmp_path / "in")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rust_crosscheck", "export_inputs.py"), d], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rust_crosscheck", "compare.py"), d, "--self-test"], capture_output=True, text=True)
    assert r.returncode == 0 and "identical to the oracle" in r.stdout, r.stdout + r.stderr
    # the dumper addresses exactly the files the exporter wrote
    src = open(os.path.join(ROOT, "tools", "rust_crosscheck", "crosscheck.rs")).read()
    for stem in ("commit_", "fri_", "perm_", "gate_"):
        assert stem in src
