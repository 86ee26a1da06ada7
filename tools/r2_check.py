#!/usr/bin/env python
"""Box visit (round 2): parity of the TMA kernel / sparse start on small registers, then per-pass times of
the headline circuit with the TMA kernel on and off."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import ref_sim
from qrisp_b200.circuit import ghz_qft_circuit, random_circuit, qft_circuit
from qrisp_b200.engine import StateVector

def sv_of(qc, **kw):
    sv = StateVector(qc.num_qubits, **kw)
    sv.apply_circuit([i for i in qc.data if i.matrix is not None])
    return sv.statevector(), sv

bad = 0
for n in (13, 14, 16, 18, 20):
    for name, qc in (("ghzqft", ghz_qft_circuit(n)), ("random", random_circuit(n, 6, seed=n)), ("qft", qft_circuit(n))):
        want = ref_sim.statevector(qc)
        for kw in (dict(tma=True, prefix_support=0), dict(tma=True), dict(tma=True, prefix_support=8), dict(tma=False), dict(tma=False, prefix_support=0)):
            try:
                got, sv = sv_of(qc, **kw)
                err = float(np.abs(got - want).max())
            except Exception as e:
                err = str(e)[:200]
            ok = isinstance(err, float) and err < 1e-5
            bad += not ok
            print(("ok  " if ok else "BAD ") + f"{name} n={n} {kw} err={err} passes={sv.stats['passes']} prefix={sv.stats.get('prefix_gates', 0)}", flush=True)
print("PARITY", "FAILED" if bad else "ok", bad, flush=True)
if bad and "--force" not in sys.argv:
    sys.exit(1)

def time_prog(n, qc, reps=5, **kw):
    sv = StateVector(n, **kw)
    instrs = [i for i in qc.data if i.matrix is not None]
    t0 = time.perf_counter()
    prog = sv.compile(instrs)
    plan_s = time.perf_counter() - t0
    for _ in range(2):
        sv.reset(); prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sv.reset(); prog.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # per pass
    sv.reset(); prog._enter()
    pev = []
    for i, s in enumerate(prog.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if s.kind == "pass":
            sv._run_pass(s, prog._headers[i], prog.dev, prog.offsets[i], sparse=prog._entry_sparse if i == 0 else None)
        else:
            sv._run_big(s)
        b.record(); pev.append((a, b))
    torch.cuda.synchronize()
    sv.pos[:] = prog.pos_out
    per = [round(a.elapsed_time(b), 3) for a, b in pev]
    nrm = sv.norm2()
    out = dict(n=n, kw=kw, ms=round(ms, 3), per_pass_ms=per, passes=prog.n_passes, phases=[s.n_phases for s in prog.steps if s.kind == "pass"],
               tma=[getattr(s, "tma", None) for s in prog.steps], plan_s=round(plan_s, 4), norm2=nrm,
               tiles=[s.tile_pos for s in prog.steps if s.kind == "pass"])
    print(json.dumps(out), flush=True)
    del sv, prog
    torch.cuda.empty_cache()

n = int(os.environ.get("R2_QUBITS", 30))
for kw in (dict(tma=True), dict(tma=True, prefix_support=2), dict(tma=True, prefix_support=0), dict(tma=False, prefix_support=0), dict(tma=False)):
    time_prog(n, ghz_qft_circuit(n), **kw)
for kw in (dict(tma=True, prefix_support=0), dict(tma=False, prefix_support=0)):
    time_prog(n, qft_circuit(n), **kw)
    time_prog(n, random_circuit(n, 20, seed=0), reps=2, **kw)
