#!/usr/bin/env python
"""BASELINE.json configs[2]: the README's find_order circuits (tests/golden/shor.npz, recorded by
tests/golden/make_golden.py make_shor) through the drop-in run() on one B200, or sharded under torchrun.

    python tools/shor_run.py shor_order_a2_N77

Prints one JSON line: wall time, number of outcomes, agreement with the reference's result when the fixture has
one, and the number-theory check (QPE peaks sit at multiples of 2**m / r, r = the multiplicative order)."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qrisp_b200.circuit import Circuit
from qrisp_b200.simulator import run

case = sys.argv[1]
g = np.load(os.path.join(ROOT, "tests", "golden", "shor.npz"))
qc = Circuit.from_arrays(g, case + "_")
a_, N = int(case.split("_a")[1].split("_")[0]), int(case.split("_N")[1])
order = next(r for r in range(1, N) if pow(a_, r, N) == 1)
group = None
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    import torch.distributed as dist
    lr = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); group = dist.group.WORLD
t0 = time.perf_counter()
res, sv = run(qc, None, return_engine=True, group=group)
torch.cuda.synchronize()
secs = time.perf_counter() - t0
m = len(next(iter(res)))
# QPE: outcome y (as printed, clbit 0 right-most) ~ k * 2**m / order
peaks = sorted(res, key=res.get, reverse=True)[:order]
dist_to_grid = [min(abs(int(k, 2) / 2.0**m * order - j) for j in range(order + 1)) for k in peaks]
mass_near = sum(p for k, p in res.items() if min(abs(int(k, 2) / 2.0**m * order - j) for j in range(order + 1)) < order / 2.0**m * 2 + 1e-9)
out = {"case": case, "gpus": world, "qubits_circuit": qc.num_qubits, "qubits_simulated": sv.num_qubits, "instructions": len(qc.data),
       "run_s": round(secs, 2), "outcomes": len(res), "order": order, "measured_bits": m,
       "top_peaks_max_distance_to_k_over_r_in_units_of_1_over_r": round(max(dist_to_grid), 4),
       "probability_within_2_grid_points_of_k_over_r": round(mass_near, 4), "stats": {k: int(v) for k, v in sv.stats.items()}}
if case + "_keys" in g.files:
    ref = {str(k): float(p) for k, p in zip(g[case + "_keys"], g[case + "_probs"])}
    out["reference_s"] = float(g[case + "_ref_seconds"])
    out["labels_equal"] = set(res) == set(ref)
    out["max_abs_prob_err"] = max(abs(res.get(k, 0.0) - ref.get(k, 0.0)) for k in set(res) | set(ref))
else:
    out["reference"] = "fails: " + str(g[case + "_ref_failure"])
if int(os.environ.get("RANK", 0)) == 0:
    print(json.dumps(out), flush=True)
