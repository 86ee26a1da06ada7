#!/usr/bin/env python
"""Interpreter vs specialised pass kernels on one circuit: per-circuit device time of the resident program.
    python tools/jit_bench.py [ghzqft|qft|random] [qubits] [depth]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qrisp_b200.circuit import ghz_qft_circuit, qft_circuit, random_circuit
from qrisp_b200.engine import StateVector

kind = sys.argv[1] if len(sys.argv) > 1 else "random"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
qc = ghz_qft_circuit(n) if kind == "ghzqft" else qft_circuit(n) if kind == "qft" else random_circuit(n, depth, seed=0)
instrs = [i for i in qc.data if i.matrix is not None]
out = {"circuit": kind, "qubits": n, "gates": len(instrs)}
ref = None
for label in ("interpreter", "specialised"):
    sv = StateVector(n, specialize=False)
    prog = sv.compile(instrs)
    if label == "specialised":
        t0 = time.perf_counter()
        out["kernels"] = prog.specialize()
        out["compile_s"] = round(time.perf_counter() - t0, 1)
    for _ in range(2):
        sv.reset(); prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        sv.reset(); prog.run()
    e1.record(); torch.cuda.synchronize()
    out[label + "_ms"] = round(e0.elapsed_time(e1) / reps, 3)
    out[label + "_norm2"] = sv.norm2()
    idx = np.random.default_rng(0).integers(0, 1 << n, 1000, dtype=np.uint64)
    amps = sv.amplitudes(idx)
    if ref is None:
        ref = amps
    else:
        out["max_abs_diff_at_1000_amplitudes"] = float(np.abs(amps - ref).max())
    out["passes"] = prog.n_passes
    del sv, prog
    torch.cuda.empty_cache()
print(json.dumps(out))
