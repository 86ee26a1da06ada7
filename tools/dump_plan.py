#!/usr/bin/env python
"""Host-only: print the plan of a circuit (passes, phases, op records) -- what the pass kernel will decode.

    python tools/dump_plan.py ghzqft 30 [prefix_support [ranks]]      (ranks: the plan of a sharded state)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pass_emulator
from qrisp_b200.circuit import ghz_qft_circuit, qft_circuit, random_circuit
from qrisp_b200.distributed import PlanOnlyShard

KIND = {1: "MAT1", 2: "MAT2", 3: "MAT3", 4: "MAT4", 5: "DIAG", 6: "DIAG_THR", 7: "DIAG_BIT", 8: "DIAG_VEC", 9: "BFLY"}
name, n = sys.argv[1], int(sys.argv[2])
ps = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-" else None
world = int(sys.argv[4]) if len(sys.argv) > 4 else 1
qc = {"ghzqft": ghz_qft_circuit, "qft": qft_circuit}.get(name, lambda n: random_circuit(n, 20, seed=0))(n)
instrs = [i for i in qc.data if i.matrix is not None]
sv = PlanOnlyShard(n, world, prefix_support=ps)
steps, entry, pos_in = sv.plan(instrs)
print(f"{len(instrs)} gates, {world} rank(s), host prefix takes {sv.stats.get('prefix_gates', 0)}, support {len(entry[0]) if entry else 0}; entry map {pos_in}")
for k, s in enumerate(steps):
    if s.kind != "pass":
        print(f"step {k}: {s.kind}" + (f" evict {s.evict} perm {s.perm}" if s.kind == "swap" else ""))
        continue
    P = pass_emulator.parse(s.blob)
    print(f"pass {k}: tile {s.tile_pos} tma={s.tma} phases={s.n_phases} records={len(P.ops)} slots={P.n_slots} small={P.small} total={len(s.blob)}")
    for pi, ph in enumerate(P.phases):
        first, cnt = ph[0], ph[1]
        print(f"  phase {pi}: xflags={ph[6] if len(ph) > 6 else None:#x}")
        for o in P.ops[first:first + cnt]:
            print(f"    {KIND.get(o['kind'], o['kind']):9s} code={o['code']:3d} b0={o['b0']} b1={o['b1']} fl={o['flags']:#04x} nthr={o['nthr']} nreg={o['nreg']} jmask={o['jmask']:#010x} cmask={o['cmask']:#x}")
