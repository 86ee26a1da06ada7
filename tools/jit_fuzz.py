#!/usr/bin/env python
"""Fuzz of the specialised kernels: random gate soups (tests/soup.py) over every tile geometry, complex64 and
complex128, specialised kernels (compiled on the spot, in parallel) against the interpreter -- bit for bit -- and
against the oracle.     python tools/jit_fuzz.py [cases]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import ref_sim
from soup import gate_soup
from qrisp_b200.engine import StateVector

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rng = np.random.default_rng(2026)
geos = [{}, {"T": 9, "r": 3}, {"T": 10, "r": 4}, {"T": 12, "r": 5}, {"T": 13, "r": 4}, {"dtype": np.complex128}, {"dtype": np.complex128, "r": 4, "T": 11}]
bad, kernels, t0 = 0, 0, time.time()
for case in range(cases):
    n = int(rng.integers(11, 17))
    g = geos[case % len(geos)]
    if g.get("T", 0) > n:
        continue
    qc = gate_soup(n, int(rng.integers(10, 70)), rng, wide=case % 2 == 1)
    instrs = [i for i in qc.data if i.matrix is not None]
    res = []
    for spec in (False, True):
        sv = StateVector(n, specialize=False, prefix_support=int(rng.integers(0, 2)) * 64 if spec is False else res[0][2], **g)
        ps = sv.prefix_support
        prog = sv.compile(instrs)
        if spec:
            kernels += prog.specialize()
        prog.run()
        res.append((sv.statevector(), sv.stats.get("jit_passes", 0), ps))
    c128 = g.get("dtype") is np.complex128
    want = ref_sim.statevector(qc, dtype=np.complex128)
    same = np.array_equal(res[0][0], res[1][0])
    err = float(np.abs(res[1][0] - want).max())
    ok = same and err < (1e-11 if c128 else 3e-5) and (res[1][1] > 0 or prog.n_passes == 0)  # (the host prefix may take everything)
    bad += not ok
    print(("ok  " if ok else "BAD ") + f"case {case} n={n} {g} gates={len(instrs)} jit_passes={res[1][1]} bit_identical={same} err={err:.2e}", flush=True)
print(f"JIT FUZZ {'FAILED' if bad else 'ok'}: {bad} bad of {cases}, {kernels} specialised passes, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
