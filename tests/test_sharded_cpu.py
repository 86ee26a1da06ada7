"""The multi-GPU path on CPU: world_size 2, 4 and 8 over gloo.

The sharding logic (swap planning, position bookkeeping, the all-to-all exchange, rank bits
in ``hi_base``) is the product code of ``qrisp_b200/distributed.py``; only the four device
hooks are replaced by CPU stand-ins (numpy state, the pass emulator from ``oracle/``), since
there is no GPU here.  The result of every rank is gathered and compared with the oracle.
"""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import big_step_numpy
        from oracle import pass_emulator, ref_sim
        from qrisp_b200.circuit import Circuit, ghz_qft_circuit, random_circuit
        from qrisp_b200.distributed import ShardedStateVector, _canonical

        class CpuShard(ShardedStateVector):
            prefix_support = 2  # a short host prefix (a GHZ ladder fits): the default (1024) would finish these circuits before any swap

            def _setup_device(self, device):
                self.device = torch.device("cpu")
                self.state = torch.zeros(2 << self.n, dtype=torch.float32)
                self._scratch = torch.zeros_like(self.state)

            # reset() is the product's: the state is a pending sparse list until the first pass runs

            def _sparse_for(self, step, state):
                from qrisp_b200 import prefix

                phys = prefix.to_physical(state[0], self._pending_map(), len(self.pos))
                data, k = prefix.sparse_input_list(step, phys, state[1], self.n, self.hi_base, np.complex64)
                pad = (4 * k + 15) // 16 * 16
                return (np.frombuffer(data, np.uint32, k, 0), np.frombuffer(data, np.uint32, k, pad),
                        np.frombuffer(data, np.complex64, k, 2 * pad)), k

            def _init_sparse(self, state):
                from qrisp_b200 import prefix

                phys = prefix.to_physical(state[0], self._pending_map(), len(self.pos))
                mine = (phys >> np.uint64(self.n)) == np.uint64(self.rank)
                st = np.zeros(1 << self.n, dtype=np.complex64)
                st[(phys[mine] & np.uint64((1 << self.n) - 1)).astype(np.int64)] = state[1][mine].astype(np.complex64)
                self.state = torch.from_numpy(st.view(np.float32).copy())

            def _np(self):
                return self.state.numpy().view(np.complex64)

            def _permute_local(self, perm):
                src = self._np()
                i = np.arange(1 << self.n)
                j = np.zeros_like(i)
                for p in range(self.n):
                    j |= ((i >> p) & 1) << perm[p]
                dst = np.empty_like(src)
                dst[j] = src
                self.state = torch.from_numpy(dst.view(np.float32).copy())

            def _upload_blobs(self, steps, offs, total):
                return None

            def _run_pass(self, step, header, dev, offset, sparse=None, norms=False):
                lst = None
                if self._pending is not None:
                    lst = (sparse if sparse is not None else self._sparse_for(step, self._pending))[0]
                    self._pending = None
                out = pass_emulator.run(step.blob, self._np(), hi_base=self.hi_base, sparse=lst)
                self.state = torch.from_numpy(out.view(np.float32).copy())

            def _run_big(self, s):
                # rank bits enter the control mask through the physical index
                self._materialize()
                st = self._np()
                idx = np.arange(st.size) | self.hi_base
                full_on = ((idx ^ s.cval) & s.cmask) == 0
                local = type("S", (), dict(positions=s.positions, matrix=s.matrix, cmask=0, cval=0))
                out = big_step_numpy(st, local)
                out[~full_on] = st[~full_on]
                self.state = torch.from_numpy(out.view(np.float32).copy())

        if case == "qft_fixed_layout":
            # no free choice of the entry map, no host prefix: qubit 0 sits on a rank bit and is the first dense
            # target, so the plan OPENS with a swap on a state nobody has written yet
            os.environ["QB2_NO_INITIAL_LAYOUT"] = "1"
            CpuShard.prefix_support = 0
            qc = ghz_qft_circuit(13)
        elif case in ("qft_long_prefix", "soup_long_prefix"):
            # the product's default: the host prefix walks `world` times further than on one device and every
            # rank builds its first tiles from its own share of the list
            CpuShard.prefix_support = None
            if case == "qft_long_prefix":
                qc = ghz_qft_circuit(13)
            else:
                from soup import gate_soup

                qc = Circuit(13)
                for q in range(13):
                    qc.h(q)
                for ins in gate_soup(13, 80, np.random.default_rng(int(os.environ.get("QB2_FUZZ_SEED", 5))), wide=True).data:
                    qc.data.append(ins)
        elif case == "qft":
            qc = ghz_qft_circuit(13)
        elif case == "soup":
            from soup import gate_soup

            qc = gate_soup(13, 120, np.random.default_rng(int(os.environ.get("QB2_FUZZ_SEED", 99))), wide=True)
        elif case == "random":
            qc = random_circuit(13, 6, seed=21, two_qubit="cx")
            qc.mcx([0, 12], 6)
            qc.cp(0.3, 0, 12)
            qc.swap(1, 11)
            qc.rxx(0.7, 0, 5)
        else:
            rng = np.random.default_rng(3)
            a = rng.normal(size=(32, 32)) + 1j * rng.normal(size=(32, 32))
            u, _ = np.linalg.qr(a)
            qc = random_circuit(13, 2, seed=5)
            qc.unitary(u, (0, 3, 7, 9, 12))  # wider than a pass: the big-matrix path, with a swap
            qc.h(0)
        instrs = [i for i in qc.data if i.matrix is not None]
        sv = CpuShard(13)
        prog = sv.compile(instrs)
        prog.run()
        sv._materialize()  # (a circuit the host prefix absorbs whole leaves the state pending)
        shards = [torch.empty_like(sv.state) for _ in range(world)]
        dist.all_gather(shards, sv.state)
        full = torch.cat(shards).numpy().view(np.complex64)
        got = _canonical(full, sv.pos, 13)
        want = ref_sim.statevector(qc)
        err = float(np.abs(got - want).max())
        n_swaps = sum(1 for s in prog.steps if s.kind == "swap")
        # the streaming path (apply_circuit: every step launched the moment it is planned) gives the same state
        sv2 = CpuShard(13)
        steps2 = sv2.apply_circuit(instrs)
        sv2._materialize()
        shards = [torch.empty_like(sv2.state) for _ in range(world)]
        dist.all_gather(shards, sv2.state)
        got2 = _canonical(torch.cat(shards).numpy().view(np.complex64), sv2.pos, 13)
        assert [s.kind for s in steps2] == [s.kind for s in prog.steps] and list(sv2.pos) == list(sv.pos)
        assert np.array_equal(got2, got), float(np.abs(got2 - got).max())
        if rank == 0:
            with open(os.path.join(out_dir, f"{case}_{world}.txt"), "w") as f:
                f.write(f"{err} {n_swaps} {sv.stats['swaps']}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "qft"), (2, "random"), (4, "qft"), (4, "random"), (2, "big"), (4, "soup"), (2, "soup"), (2, "qft_fixed_layout"),
                                        (4, "qft_long_prefix"), (2, "soup_long_prefix"), (8, "qft_long_prefix"), (8, "random")])
def test_sharded_statevector_over_gloo(tmp_path, world, case):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    err, planned, ran = open(tmp_path / f"{case}_{world}.txt").read().split()
    assert float(err) < 2e-6, err
    assert int(planned) == int(ran)
    if "long_prefix" not in case:
        assert int(planned) >= 1  # rank bits were dense targets: swaps happened


def test_eviction_choice_prefers_idle_positions():
    sys.path.insert(0, ROOT)
    from qrisp_b200.distributed import choose_evictions
    from qrisp_b200.normalize import CMat

    pos = [9 - q for q in range(10)]  # qubit q at position 9-q; 8 local positions, 2 rank bits
    h = np.eye(2)
    pending = [CMat([0], h), CMat([2], h), CMat([3], h), CMat([9], h)]
    ev = choose_evictions(pending, pos, 8, 2, protect={0, 1, 2, 3})
    # positions 7 (qubit 2) and 6 (qubit 3) are needed soon, 0..3 are protected: evict 4 and 5
    assert ev == [4, 5]


def test_initial_layout_halves_the_swaps_of_ghz_qft():
    """A fresh |0...0> does not care which qubit sits on the rank bits: the sharded engine starts
    with the qubits that are needed last up there (planning only, no process group needed)."""
    from qrisp_b200 import distributed as D
    from qrisp_b200.circuit import ghz_qft_circuit
    from qrisp_b200.normalize import normalize_circuit
    from qrisp_b200.planner import PlanConfig, Planner

    def swaps(nq, g, choose):
        class Shard:
            pass

        sv = Shard()
        sv.g, sv.n, sv.n_total = g, nq - g, nq
        sv.cfg = PlanConfig(sv.n, T=9, r=3)
        sv.pos = [nq - 1 - q for q in range(nq)]
        ops = normalize_circuit([i for i in ghz_qft_circuit(nq).data if i.matrix is not None])
        if choose:
            D.ShardedStateVector._choose_initial_layout(sv, ops)
            assert sorted(sv.pos) == list(range(nq))
        planner = Planner(sv.cfg, sv.pos, n_total=nq)
        pending, count = list(ops), 0
        while pending:
            _, need = planner.plan(pending)
            if need is None:
                break
            pending = planner.pending
            D.ShardedStateVector._plan_swap(sv, pending, set(range(sv.cfg.L)))
            count += 1
        return count

    assert swaps(20, 2, False) == 2
    assert swaps(20, 2, True) == 1
