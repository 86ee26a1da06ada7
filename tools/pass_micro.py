#!/usr/bin/env python
"""Micro-benchmarks of the pass kernel: device time of single passes with a known op mix,
so that the cost of each kernel feature (skeleton, exchange, op kinds) is measured on its
own.  Usage (GPU box):  python tools/pass_micro.py [n_qubits] [case ...]
"""

import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from qrisp_b200.circuit import Circuit
from qrisp_b200.engine import StateVector


def cases(n):
    """name -> circuit.  Qubit q sits at position n-1-q: qubits 0..3 are the top positions
    (register bits of a default phase), qubit n-1 is position 0 (a lane bit)."""
    out = {}
    hi = [0, 1, 2, 3]  # top four positions -> register bits of the first phase

    def circ(f):
        qc = Circuit(n)
        f(qc)
        return qc

    out["copy"] = circ(lambda qc: qc.p(0.0 + 1e-300, 0))  # placeholder replaced below
    out["h4"] = circ(lambda qc: [qc.h(q) for q in hi])
    out["h8"] = circ(lambda qc: [qc.h(q) for q in hi] + [qc.rz(0.3, q) for q in hi] + [qc.h(q) for q in hi])
    out["u3x4"] = circ(lambda qc: [qc.u3(0.3, 0.2, 0.1, q) for q in hi])
    out["u3x16"] = circ(lambda qc: [qc.u3(0.3 + 0.1 * k, 0.2, 0.1, q) or qc.cz(q, (q + 1) % 4) for k in range(4) for q in hi])
    for reps in (2, 4, 8):  # slope = cost per op once a single-phase pass is compute bound
        out[f"u3x{4 * reps}_1phase"] = circ(lambda qc, reps=reps: [qc.u3(0.3 + 0.1 * k, 0.2, 0.1, q) or qc.cx(q, (q + 1) % 4) for k in range(reps) for q in hi])
        out[f"bfly{4 * reps}_1phase"] = circ(lambda qc, reps=reps: [(qc.h(q), [qc.cp(math.pi / (1 << (k - q)), k, q) for k in range(q + 1, n)]) for _ in range(reps) for q in hi])
    out["cx4_thread_ctrl"] = circ(lambda qc: [qc.cx(n - 1, q) for q in hi])
    out["cx4_reg_ctrl"] = circ(lambda qc: [qc.cx(hi[(i + 1) % 4], hi[i]) for i in range(4)])
    out["cp_ladder4"] = circ(lambda qc: [qc.cp(math.pi / (1 << (k - q)), k, q) for q in hi for k in range(q + 1, n)])
    out["h_cp_ladder4"] = circ(lambda qc: [(qc.h(q), [qc.cp(math.pi / (1 << (k - q)), k, q) for k in range(q + 1, n)]) for q in hi])
    out["cz_layer"] = circ(lambda qc: [qc.cz(2 * i, 2 * i + 1) for i in range(n // 2)])
    out["h8_two_phases"] = circ(lambda qc: [qc.h(q) for q in range(8)])
    out["h12_three_phases"] = circ(lambda qc: [qc.h(q) for q in range(8)] + [qc.h(n - 1 - q) for q in range(4)])
    out["rxx2"] = circ(lambda qc: [qc.rxx(0.3, 0, 1), qc.rxx(0.4, 2, 3)])
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 28
    want = [a for a in sys.argv[1:] if not a.isdigit()]
    sv = StateVector(n)
    sv.apply_circuit([i for i in cases(n)["h4"].data])  # a non-trivial state
    state_bytes = 8 << n
    res = {}
    for name, qc in cases(n).items():
        if want and name not in want:
            continue
        pos0 = list(sv.pos)
        instrs = [i for i in qc.data if i.matrix is not None]
        if name == "copy":
            instrs = []
        prog = sv.compile(instrs) if instrs else None
        if prog is None:
            # an empty pass: plan a single identity-free op list by hand
            from qrisp_b200.planner import Planner

            pl = Planner(sv.cfg, sv.pos)
            step, _sigma = pl._build_pass([("diag", (sv.cfg.n - 1,), np.array([0.0, 0.0]))], set())
            from qrisp_b200.engine import Program

            prog = Program(sv, [step])
        sv._materialize()
        def once():
            sv.pos[:] = prog.pos_in  # back-to-back replays of one pass: only the time matters
            prog.run()

        for _ in range(int(os.environ.get('MICRO_WARM', 2))):
            once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = int(os.environ.get('MICRO_REPS', 5))
        e0.record()
        for _ in range(reps):
            once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        sv.pos[:] = pos0
        res[name] = {
            "ms": round(ms, 4),
            "passes": prog.n_passes,
            "phases": prog.n_phases,
            "kernel_ops": prog.n_kernel_ops,
            "GBps_per_pass": round(2 * state_bytes * prog.n_passes / (ms * 1e-3) / 1e9, 1),
        }
        print(name, json.dumps(res[name]), flush=True)
    return res


if __name__ == "__main__":
    main()
