import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from qrisp_b200.circuit import ghz_qft_circuit, Circuit
from qrisp_b200 import simulator, engine
from qrisp_b200.preprocess import unitarize
n = 30
qc = ghz_qft_circuit(n); qc.measure(list(range(n)))
simulator.sample_counts(qc, 16, seed=0); torch.cuda.synchronize()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    instrs, nn, meas = unitarize(qc); t1 = time.perf_counter()
    sv = engine.StateVector(nn); t2 = time.perf_counter()
    sv.apply_circuit(instrs); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    meas.sort(key=lambda x: -x[1]); mq = [q for q, _ in meas]
    outs = sv.sample(4096, mq, rng=np.random.default_rng(rep)); t5 = time.perf_counter()
    vals, counts = np.unique(outs, return_counts=True)
    d = {bin(int(v))[2:].zfill(len(mq)): int(c) for v, c in zip(vals, counts)}; t6 = time.perf_counter()
    print("unitarize %.2f  ctor %.2f  plan+launch %.2f  wait-for-device %.2f  sample %.2f  dict %.2f  total %.2f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0)))
    del sv
