"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank compares the gathered sharded state with the oracle (tolerance 1e-5, complex64),
then checks marginal probabilities and device-side sampling across shards.
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_sim  # noqa: E402
from qrisp_b200.circuit import ghz_circuit, ghz_qft_circuit, random_circuit  # noqa: E402
from qrisp_b200.distributed import ShardedStateVector  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    for name, qc in (
        ("ghzqft18", ghz_qft_circuit(18)),
        ("random17", random_circuit(17, 6, seed=2, two_qubit="cx")),
    ):
        sv = ShardedStateVector(qc.num_qubits)
        prog = sv.apply_circuit([i for i in qc.data if i.matrix is not None])
        got = sv.statevector()
        want = ref_sim.statevector(qc)
        err = float(np.abs(got - want).max())
        worst = max(worst, err)
        assert err < 1e-5, (name, err)
        assert prog.n_swaps >= 1
        assert abs(sv.norm2() - 1.0) < 1e-5
        qubits = [0, qc.num_qubits - 1, 5]
        p = sv.probabilities(qubits)
        np.testing.assert_allclose(p, ref_sim.marginal_probabilities(want.astype(np.complex128), qubits, qc.num_qubits), atol=1e-5)
        if rank == 0:
            print(f"{name}: max |amp - oracle| = {err:.2e}, swaps = {prog.n_swaps}, passes = {prog.n_passes}")
    # sampling across shards: a 24-qubit GHZ state has its two basis states on different ranks
    n = 24
    sv = ShardedStateVector(n)
    sv.apply_circuit(ghz_circuit(n).data)
    outs = sv.sample(4000, list(range(n)), rng=np.random.default_rng(7))
    vals, counts = np.unique(outs, return_counts=True)
    assert set(int(v) for v in vals) == {0, (1 << n) - 1}, vals
    assert abs(int(counts[0]) - 2000) < 200
    if rank == 0:
        print(f"dist_gpu_check ok on {world} GPUs, worst error {worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
