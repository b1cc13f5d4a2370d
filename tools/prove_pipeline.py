"""The data path of prove() (plonky2/src/plonk/prover.rs:239-700 `my_prove`) end to end on the device, on a synthetic
witness: everything between "the witness matrix exists" and "the proof's field elements exist", i.e.

  commit wires -> Z / partial products -> commit them -> quotient values + coefficients -> commit the quotient chunks ->
  openings at zeta, g*zeta -> FRI opening proof (commit phase, PoW, query rounds)

with the transcript steps between the calls done by a stand-in (fixed challenges: the work does not depend on their values).
Not included (control plane, out of scope): circuit building, witness generation, the preprocessed constants_sigmas
commit (done once per circuit; built here outside the timed region like the reference's CircuitData).

  python tools/prove_pipeline.py ecc 17        # BASELINE config 4 shape: 2^17 rows, wide_ecc_config (234 wires), U32-heavy gates
  python tools/prove_pipeline.py recursion 16  # BASELINE config 3 shape: standard_recursion_config (135 wires), recursion gates
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky2_gpu_b200 as p2b
from plonky2_gpu_b200.pipeline import ProvePipeline

kind = sys.argv[1] if len(sys.argv) > 1 else "ecc"
n_log = int(sys.argv[2]) if len(sys.argv) > 2 else 17
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = p2b.Context(0)
pipe = ProvePipeline(ctx, kind, n_log)
pipe.prove({})
times, walls = {}, []
for _ in range(reps):
    walls.append(pipe.prove(times))
print("prove() data path, " + pipe.describe())
total = 0.0
for k, v in times.items():
    print("  %-26s %9.2f ms" % (k, min(v)))
    total += min(v)
print("  %-26s %9.2f ms (sum of stage minima; host wall per proof %.2f ms, min of %d)" % ("total", total, min(walls), reps))
