#!/usr/bin/env python
"""Host-only: print the plan of a circuit (passes, phases, op records) -- what the pass kernel will decode.

    python tools/dump_plan.py ghzqft 30 [prefix_support]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pass_emulator
from qrisp_b200 import prefix
from qrisp_b200.circuit import ghz_qft_circuit, qft_circuit, random_circuit
from qrisp_b200.normalize import normalize_circuit
from qrisp_b200.planner import PlanConfig, Planner

KIND = {1: "MAT1", 2: "MAT2", 3: "MAT3", 4: "MAT4", 5: "DIAG", 6: "DIAG_THR", 7: "DIAG_BIT", 8: "DIAG_VEC", 9: "BFLY"}
name, n = sys.argv[1], int(sys.argv[2])
ps = int(sys.argv[3]) if len(sys.argv) > 3 else None
qc = {"ghzqft": ghz_qft_circuit, "qft": qft_circuit}.get(name, lambda n: random_circuit(n, 20, seed=0))(n)
instrs = [i for i in qc.data if i.matrix is not None]
consumed, state = prefix.sparse_prefix(instrs, prefix.zero_state(), ps)
print(f"{len(instrs)} gates, host prefix takes {consumed}, support {len(state[0])}")
cfg = PlanConfig(n)
pos = [n - 1 - q for q in range(n)]
ops = normalize_circuit(instrs[consumed:], max_targets=cfg.max_mat, unit_tol=1e-6)
steps, need = Planner(cfg, pos).plan(ops)
for k, s in enumerate(steps):
    if s.kind != "pass":
        print(f"step {k}: {s.kind}")
        continue
    P = pass_emulator.parse(s.blob)
    print(f"pass {k}: tile {s.tile_pos} tma={s.tma} phases={s.n_phases} records={len(P.ops)} slots={P.n_slots} small={P.small} total={len(s.blob)}")
    for pi, ph in enumerate(P.phases):
        first, cnt = ph[0], ph[1]
        print(f"  phase {pi}: xflags={ph[6] if len(ph) > 6 else None:#x}")
        for o in P.ops[first:first + cnt]:
            print(f"    {KIND.get(o['kind'], o['kind']):9s} code={o['code']:3d} b0={o['b0']} b1={o['b1']} fl={o['flags']:#04x} nthr={o['nthr']} nreg={o['nreg']} jmask={o['jmask']:#010x} cmask={o['cmask']:#x}")
