"""Parity of the CUDA path against the oracle on small registers, single GPU or sharded:

    python tests/dist_gpu_check.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py                  # one rank per GPU

``parity_cases(group)`` is also what ``bench.py`` runs, untimed, before its timed region (the driver's
scaling run then carries a parity record for every N).  Every rank compares the gathered state with the
oracle (tolerance 1e-5, complex64), then marginal probabilities, then device-side sampling.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parity_cases(group=None, verbose=False):
    """Runs the cases on the current device (``group`` None) or sharded over ``group``.  Returns
    ``{"cases": [...], "max_abs_err": e, "swaps": k}``; raises AssertionError on a mismatch."""
    from oracle import ref_sim  # the checker: never the thing measured
    from qrisp_b200.circuit import ghz_circuit, ghz_qft_circuit, random_circuit
    from qrisp_b200.engine import StateVector

    if group is not None:
        import torch.distributed as dist

        from qrisp_b200.distributed import ShardedStateVector

        world, rank = dist.get_world_size(group), dist.get_rank(group)

        def make(n, **kw):
            return ShardedStateVector(n, group=group, **kw)
    else:
        world, rank = 1, 0

        def make(n, **kw):
            return StateVector(n, **kw)

    worst, swaps, names = 0.0, 0, []
    for name, qc, kw in (
        ("ghzqft18", ghz_qft_circuit(18), {}),
        ("ghzqft18_short_prefix", ghz_qft_circuit(18), {"prefix_support": 2}),
        ("random17", random_circuit(17, 6, seed=2, two_qubit="cx"), {}),
        ("random17_no_tma", random_circuit(17, 6, seed=2, two_qubit="cx"), {"tma": False, "prefix_support": 0}),
    ):
        sv = make(qc.num_qubits, **kw)
        sv.apply_circuit([i for i in qc.data if i.matrix is not None])
        got = sv.statevector()
        want = ref_sim.statevector(qc)
        err = float(np.abs(got - want).max())
        worst = max(worst, err)
        assert err < 1e-5, (name, err)
        assert abs(sv.norm2() - 1.0) < 1e-5
        qubits = [0, qc.num_qubits - 1, 5]
        p = sv.probabilities(qubits)
        np.testing.assert_allclose(p, ref_sim.marginal_probabilities(want.astype(np.complex128), qubits, qc.num_qubits), atol=1e-5)
        swaps += sv.stats.get("swaps", 0)
        names.append(name)
        if verbose and rank == 0:
            print(f"{name}: max |amp - oracle| = {err:.2e}, swaps = {sv.stats.get('swaps', 0)}, passes = {sv.stats['passes']}")
    if world > 1:
        assert swaps >= 1, "no qubit swap was exercised"
    # sampling (across shards: a 24-qubit GHZ state has its two basis states on different ranks)
    n = 24
    sv = make(n)
    sv.apply_circuit(ghz_circuit(n).data)
    outs = sv.sample(4000, list(range(n)), rng=np.random.default_rng(7))
    vals, counts = np.unique(outs, return_counts=True)
    assert set(int(v) for v in vals) == {0, (1 << n) - 1}, vals
    assert abs(int(counts[0]) - 2000) < 200
    names.append("ghz24_sampling")
    return {"cases": names, "max_abs_err": worst, "swaps": int(swaps), "tolerance": 1e-5, "against": "oracle/ref_sim.py"}


def main():
    import torch

    world = int(os.environ.get("WORLD_SIZE", 1))
    group = None
    if world > 1:
        import torch.distributed as dist

        local_rank = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    res = parity_cases(group, verbose=True)
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"dist_gpu_check ok on {world} GPU(s), worst error {res['max_abs_err']:.2e}, {res['swaps']} qubit swaps")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
