#!/usr/bin/env python
"""Where the end-to-end time of run(qc, 4096) goes (GHZ+QFT, 30 qubits): host steps with a device
synchronisation after each (so the sum is LARGER than the pipelined call, which is printed beside it)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qrisp_b200.circuit import ghz_qft_circuit
from qrisp_b200 import simulator, engine
from qrisp_b200.normalize import normalize_circuit
from qrisp_b200.planner import Planner
from qrisp_b200.preprocess import unitarize

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
qc = ghz_qft_circuit(n); qc.measure(list(range(n)))
simulator.run(qc, 16, seed=0); torch.cuda.synchronize()
sync = torch.cuda.synchronize
for rep in range(3):
    sync(); t0 = time.perf_counter()
    simulator.run(qc, 4096, seed=rep); sync()
    whole = time.perf_counter() - t0
    t = [time.perf_counter()]
    def mark(): sync(); t.append(time.perf_counter())
    instrs, nn, meas = unitarize(qc); mark()
    sv = engine.StateVector(nn); mark()
    rest = sv._take_prefix(list(instrs)); mark()
    ops = normalize_circuit(rest, max_targets=sv.cfg.max_mat, unit_tol=1e-6); mark()
    steps, need = Planner(sv.cfg, list(sv.pos)).plan(ops); mark()
    sv2 = engine.StateVector(nn); sv2.apply_circuit(instrs); mark()
    meas.sort(key=lambda x: -x[1]); mq = [q for q, _ in meas]
    outs = sv2.sample(4096, mq, rng=np.random.default_rng(rep)); mark()
    d = simulator.counts_dict(outs, len(mq)); mark()
    names = ["unitarize", "ctor", "host prefix", "normalize", "plan (alone)", "ctor+apply_circuit (streamed, synced)", "sample", "counts dict"]
    print("run() %.2f ms | " % (1e3 * whole) + "  ".join("%s %.2f" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t, t[1:])), flush=True)
    del sv, sv2
