#!/usr/bin/env python
"""Dense 5-qubit blocks: the tensor-core kernel (qb2_apply_block_tc) against the SIMT path
(qb2_apply_matrix_big) on one state.     python tools/block_bench.py [qubits]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qrisp_b200.circuit import Circuit
from qrisp_b200.engine import StateVector
from qrisp_b200.planner import BigMatStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(0)
a = rng.normal(size=(32, 32)) + 1j * rng.normal(size=(32, 32))
u, _ = np.linalg.qr(a)
sv = StateVector(n)
qc = Circuit(n)
for q in range(n):
    qc.h(q)
sv.apply_circuit(qc.data)
peak = 6541.5
for name, targets in (("low", [0, 1, 2, 3, 4]), ("mid", [10, 11, 12, 13, 14]), ("high", [n - 5, n - 4, n - 3, n - 2, n - 1]),
                      ("spread", [2, 9, 15, 21, n - 1]), ("lane-ish", [5, 6, 7, 8, 9])):
    step = BigMatStep(targets, u, 0, 0)
    out = {"n": n, "targets": name}
    for label, tc in (("tc", True), ("simt", False)):
        sv.block_tc = tc
        for _ in range(2):
            sv._run_big(step)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if tc else 2
        e0.record()
        for _ in range(reps):
            sv._run_big(step)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[label + "_ms"] = round(ms, 3)
        out[label + "_frac_of_hbm_peak"] = round((16 << n) / (ms * 1e-3) / 1e9 / peak, 3)
    out["norm2"] = sv.norm2()
    print(json.dumps(out), flush=True)

# wider blocks: cosine-sine decomposition into controlled 5-qubit tensor-core blocks + multiplexed rotations
for k in (6, 7):
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    uk, _ = np.linalg.qr(a)
    targets = list(range(10, 10 + k))
    step = BigMatStep(targets, uk, 0, 0)
    out = {"n": n, "k": k, "targets": targets}
    for label, tc in (("tc_csd", True), ("simt", False)):
        sv.block_tc = tc
        sv._run_big(step)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if tc else 1
        e0.record()
        for _ in range(reps):
            sv._run_big(step)
        e1.record()
        torch.cuda.synchronize()
        out[label + "_ms"] = round(e0.elapsed_time(e1) / reps, 3)
    out["norm2"] = sv.norm2()
    print(json.dumps(out), flush=True)
