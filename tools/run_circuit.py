#!/usr/bin/env python
"""Time one synthetic circuit on one GPU:  python tools/run_circuit.py random 33 20 | qft 30"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qrisp_b200.circuit import ghz_qft_circuit, random_circuit, qft_circuit
from qrisp_b200.engine import StateVector

kind, n = sys.argv[1], int(sys.argv[2])
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
qc = random_circuit(n, depth, seed=0) if kind == "random" else (ghz_qft_circuit(n) if kind == "qft" else qft_circuit(n))
instrs = [i for i in qc.data if i.matrix is not None]
sv = StateVector(n)
t0 = time.perf_counter()
prog = sv.compile(instrs)
t_plan = time.perf_counter() - t0
best = None
for rep in range(int(os.environ.get("RUN_REPS", 3))):  # the first run also pays for the first touch of the buffers
    sv.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prog.run(); e1.record(); torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1)
    best = ms_ if best is None else min(best, ms_)
ms = best
norm = sv.norm2()
state_bytes = 8 << n
print(json.dumps({"circuit": kind, "qubits": n, "gates": len(instrs), "plan_s": round(t_plan, 3), "sweep_ms": round(ms, 2),
                  "gates_per_s": round(len(instrs) / (ms * 1e-3), 1), "passes": prog.n_passes, "phases": prog.n_phases,
                  "kernel_ops": prog.n_kernel_ops, "avg_pass_ms": round(ms / prog.n_passes, 3),
                  "avg_pass_GBps": round(2 * state_bytes * prog.n_passes / (ms * 1e-3) / 1e9, 1), "norm2": norm}))
