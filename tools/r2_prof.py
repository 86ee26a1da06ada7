#!/usr/bin/env python
"""Profiling target: `steps` replays of one compiled circuit (default GHZ+QFT on 30 qubits), nothing else.
    python tools/r2_prof.py [ghzqft|qft|random] [qubits] [steps]      (R2_TMA=0 / R2_PREFIX=k change the engine)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qrisp_b200.circuit import ghz_qft_circuit, qft_circuit, random_circuit
from qrisp_b200.engine import StateVector

kind = sys.argv[1] if len(sys.argv) > 1 else "ghzqft"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
qc = ghz_qft_circuit(n) if kind == "ghzqft" else qft_circuit(n) if kind == "qft" else random_circuit(n, 20, seed=0)
kw = {}
if os.environ.get("R2_TMA") is not None:
    kw["tma"] = os.environ["R2_TMA"] != "0"
if os.environ.get("R2_PREFIX") is not None:
    kw["prefix_support"] = int(os.environ["R2_PREFIX"])
sv = StateVector(n, **kw)
prog = sv.compile([i for i in qc.data if i.matrix is not None])
for _ in range(steps):
    sv.reset()
    prog.run()
torch.cuda.synchronize()
print("done", prog.n_passes, "passes per step; norm2", sv.norm2())
