#!/usr/bin/env python
"""cProfile of the host side of run(qc, 4096) (GHZ+QFT, 30 qubits): where the end-to-end time beyond the
device time goes.     python tools/e2e_profile.py [qubits]   (QB2_SPECIALIZE=1 for the specialised kernels)"""
import cProfile, io, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qrisp_b200.circuit import ghz_qft_circuit
from qrisp_b200 import simulator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
qc = ghz_qft_circuit(n); qc.measure(list(range(n)))
for _ in range(2):
    simulator.run(qc, 16, seed=0)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    simulator.run(qc, 4096, seed=i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
