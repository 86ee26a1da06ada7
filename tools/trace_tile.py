#!/usr/bin/env python
"""Where does a tile's time go?  Runs one circuit through the TRACE build of the pass kernel
(`make trace`) and prints, for CTA 0 / thread 0, the clock-cycle deltas between the stations of
its first tiles: load issue, slot sync, every op, every exchange, store.

    python tools/trace_tile.py qft 30 [pass_index]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from qrisp_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "qrisp_b200", "lib", "libqb2_trace.so")
import torch
from qrisp_b200.circuit import ghz_qft_circuit, random_circuit
from qrisp_b200.engine import StateVector

kind, n = sys.argv[1], int(sys.argv[2])
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
qc = ghz_qft_circuit(n) if kind == "qft" else random_circuit(n, 20, seed=0)
instrs = [i for i in qc.data if i.matrix is not None]
sv = StateVector(n)
prog = sv.compile(instrs)
lib = _lib.load()
lib.qb2_debug_read_trace.restype = ctypes.c_int
lib.qb2_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
TILES, STAMPS = 24, 64
buf = np.zeros(TILES * STAMPS, dtype=np.uint64)
prog.run(); torch.cuda.synchronize()
lib.qb2_debug_read_trace(buf.ctypes.data, buf.size)  # warm-up run discarded
names = {0: "start", 1: "ld_issued", 2: "slots", 62: "st_begin", 63: "st_issued"}
for p in range(1, 8):
    names[4 + 2 * p] = f"x{p}_begin"; names[5 + 2 * p] = f"x{p}_end"
for o in range(37):
    names[24 + o] = f"op{o}"
def run_one(i):
    s, o, h = prog.steps[i], prog.offsets[i], prog._headers[i]
    if s.kind == "pass":
        sv._run_pass(s, h, prog.dev, o)
    else:
        sv._run_big(s)


n_pass = len(prog.steps)
for i in range(n_pass):
    if which >= 0 and i != which:
        run_one(i); continue
    run_one(i); torch.cuda.synchronize()
    lib.qb2_debug_read_trace(buf.ctypes.data, buf.size)
    t = buf.reshape(TILES, STAMPS).astype(np.int64)
    print(f"=== pass {i}")
    for tile in (2, 3, 10):
        row = t[tile]
        idx = [k for k in sorted(names, key=lambda k: (row[k] if row[k] else 1 << 62)) if row[k]]
        if not idx:
            continue
        base = row[idx[0]]
        prev = base
        out = []
        for k in idx:
            out.append(f"{names[k]}+{row[k]-prev}")
            prev = row[k]
        nxt = t[tile + 1][0] - base if t[tile + 1][0] else 0
        print(f" tile {tile}: total {row[idx[-1]]-base} (to next start {nxt}): " + " ".join(out))
