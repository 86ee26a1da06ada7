"""Statevector sharded over several GPUs by its high physical positions.

``world = 2**g`` ranks (one process per GPU, ``torch.distributed``); rank ``k`` owns the
``2**n_local`` amplitudes whose positions ``>= n_local`` spell ``k`` (the "rank bits").

* Diagonal gates and controls on rank bits need no communication: the pass kernel receives
  the rank bits through ``hi_base`` and evaluates phases / control masks on the full
  physical index.
* A dense target on a rank bit triggers a QUBIT SWAP: ``g`` local positions (chosen by
  furthest next use) trade places with the ``g`` rank bits.  Data movement = one local bit
  permutation that brings the evicted positions to the top of the local index, then one
  ``all_to_all_single`` of ``2**g`` contiguous chunks over NCCL / NVLink.  Afterwards only
  the position map changes.

The reference has no counterpart (its simulator is single-process numpy); this is the
"statevector too large for one GPU" row of the north star.  The class mirrors
:class:`qrisp_b200.engine.StateVector` so that ``run`` / ``bench.py`` treat both alike.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib, prefix
from .engine import MIN_QUBITS, DeviceOps, Program
from .normalize import normalize_circuit
from .planner import QB2_C64, QB2_C128, PlanConfig, Planner


class SwapStep:
    """Trade the rank bits with the local positions ``evict`` (ascending list of g positions)."""

    kind = "swap"

    def __init__(self, evict, perm):
        self.evict = evict  # local positions that become rank bits
        self.perm = perm  # local bit permutation applied first (None: evict already on top)


def choose_evictions(pending, pos, n_local, g, protect):
    """The g local positions whose next use as a dense target is furthest away (Belady);
    positions in ``protect`` (the low, coalescing positions) are never evicted."""
    where = {p: q for q, p in enumerate(pos)}
    next_use = {p: len(pending) + 1 for p in range(n_local) if p not in protect}
    for i, op in enumerate(pending):
        if op.kind == "relabel":
            break  # the map changes from here on; keep what we know
        for q in op.nd:
            p = pos[q]
            if p in next_use and next_use[p] > i:
                next_use[p] = i
    order = sorted(next_use, key=lambda p: (-next_use[p], -p))
    return sorted(order[:g])


class ShardedStateVector(DeviceOps):
    def __init__(self, num_qubits, group=None, dtype=np.complex64, device=None, T=None, r=None, L=4, tma=None, prefix_support=None,
                 specialize=None):
        import torch
        import torch.distributed as dist

        from .engine import _specialize_mode

        self.specialize = _specialize_mode(specialize)

        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self._init_host(num_qubits, dist.get_world_size(self.group), dist.get_rank(self.group), dtype, T, r, L, tma, prefix_support)
        self._setup_device(device)
        self.reset()

    def _init_host(self, num_qubits, world, rank, dtype, T, r, L, tma, prefix_support):
        """Everything the planning side needs (no device, no process group)."""
        self.world, self.rank = int(world), int(rank)
        g = self.world.bit_length() - 1
        if 1 << g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.g = g
        self.num_qubits = int(num_qubits)
        self.n_total = max(self.num_qubits, MIN_QUBITS + g)
        self.n = self.n_total - g  # local positions
        if self.n < 2 * g + L:
            raise ValueError("too few local qubits for this many ranks")
        dtype = np.dtype(dtype)
        if dtype == np.complex64:
            self.dtype, self.np_dtype = QB2_C64, np.complex64
        elif dtype == np.complex128:
            self.dtype, self.np_dtype = QB2_C128, np.complex128
        else:
            raise TypeError(f"unsupported amplitude dtype {dtype}")
        self.cfg = PlanConfig(self.n, self.dtype, T=T, r=r, L=L, tma=tma)
        if prefix_support is not None:
            self.prefix_support = prefix_support
        nq = self.num_qubits
        self.pos = [nq - 1 - q for q in range(nq)] + list(range(nq, self.n_total))
        self.hi_base = self.rank << self.n
        self.stats = {"passes": 0, "kernel_ops": 0, "gate_ops": 0, "phases": 0, "big": 0, "swaps": 0}
        self.io = {"h2d": 0, "d2h": 0}

    # ------------------------------------------------------------------ device hooks
    # (tests replace these with CPU stand-ins to exercise the sharding logic over gloo)
    def _setup_device(self, device):
        import torch

        if not torch.cuda.is_available():
            raise _lib.Qb2Error("the B200 engine needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        real = torch.float32 if self.dtype == QB2_C64 else torch.float64
        self.state = torch.empty(2 << self.n, dtype=real, device=self.device)
        self._scratch = torch.empty_like(self.state)

    def _scratch_buf(self):
        return self._scratch

    def reset(self):
        """Back to |0...0>.  Lazy, like :meth:`StateVector.reset`: the state is a pending list of
        non-zero amplitudes on the host; the first pass builds its tiles from this rank's entries,
        anything else that looks at the shard first writes them.  While the state is pending no data
        sits in HBM, so the qubit map is still free (see :meth:`_choose_initial_layout`)."""
        self._pending = prefix.zero_state()
        nq = self.num_qubits
        self.pos[:] = [nq - 1 - q for q in range(nq)] + list(range(nq, self.n_total))
        if not getattr(self, "lazy_reset", True):
            self._materialize()

    def _permute_local(self, perm):
        self._materialize()
        arr = (ctypes.c_int * self.n)(*perm)
        _lib.check(
            self.lib.qb2_permute_bits(self.state.data_ptr(), self._scratch.data_ptr(), self.n, arr, self.dtype, self._stream()),
            "qb2_permute_bits",
        )
        self.state, self._scratch = self._scratch, self.state

    def _all_to_all(self):
        """state (2**g contiguous chunks) -> scratch: chunk c goes to rank c."""
        self._materialize()
        self.dist.all_to_all_single(self._scratch, self.state, group=self.group)
        self.state, self._scratch = self._scratch, self.state

    def synchronize(self):
        import torch

        torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ peer memory (NVLink / NVSwitch)
    def _peer_setup(self):
        """Map every rank's two shard buffers into this process (CUDA IPC through torch's storage sharing)
        so that a qubit swap can pull straight out of the peers' HBM (qb2_swap_pull).  Returns False when
        that is not possible here (no peer access, QB2_NO_PEER_SWAP set): the NCCL path takes over."""
        if getattr(self, "_peers", None) is not None:
            return bool(self._peers)
        self._peers = {}
        if os.environ.get("QB2_NO_PEER_SWAP") is not None or self.device.type != "cuda":
            return False
        import torch

        try:
            mine = [self.state.untyped_storage()._share_cuda_(), self._scratch.untyped_storage()._share_cuda_()]
            ptrs = [self.state.data_ptr(), self._scratch.data_ptr()]
            everyone = [None] * self.world
            self.dist.all_gather_object(everyone, (self.rank, mine), group=self.group)
            keep, table = [], {0: [None] * self.world, 1: [None] * self.world}
            for r, infos in everyone:
                for k, info in enumerate(infos):
                    if r == self.rank:
                        table[k][r] = ptrs[k]
                        continue
                    # opened with THIS rank's device current: cudaIpcOpenMemHandle then maps the peer's memory
                    # for this device (lazy peer access), which is what the pulling kernel needs
                    st = torch.UntypedStorage._new_shared_cuda(self.device.index, *info[1:])
                    keep.append(st)
                    table[k][r] = st.data_ptr()
            ok = torch.tensor([1], device=self.device)
            self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
            self._peer_keep = keep
            self._peers = {ptrs[0]: table[0], ptrs[1]: table[1]}
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        except Exception as exc:  # noqa: BLE001 -- any failure means "use NCCL", and says so once
            import warnings

            warnings.warn(f"qrisp_b200: peer-memory qubit swaps unavailable ({exc!r}); using NCCL all_to_all")
            self._peers = {}
        return bool(self._peers)

    def release_peers(self):
        """Drop the mappings of the peers' shards (they keep the peers' buffers alive) and let torch's IPC
        bookkeeping free what nobody uses any more.  Collective in spirit: every rank should call it (or drop
        its engine) at the same point."""
        if getattr(self, "_peer_keep", None):
            import torch

            torch.cuda.current_stream(self.device).synchronize()
            self._peer_keep = []
            self._peers = None
            torch.cuda.ipc_collect()

    def __del__(self):
        try:
            self.release_peers()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    def _stream_barrier(self):
        """Every rank's stream has reached this point when the tiny all-reduce completes (stream ordered)."""
        self.dist.all_reduce(self._flag, group=self.group)

    def _swap_pull(self, perm):
        """Local permutation + all-to-all of a qubit swap in one kernel over peer memory."""
        self._materialize()
        table = self._peers[self.state.data_ptr()]
        peers = (ctypes.c_void_p * self.world)(*table)
        arr = (ctypes.c_int * self.n)(*perm) if perm is not None else None
        self._stream_barrier()  # the peers' shards are complete
        _lib.check(
            self.lib.qb2_swap_pull(self._scratch.data_ptr(), peers, self.world, self.rank, self.n, arr, self.dtype, self._stream()),
            "qb2_swap_pull",
        )
        self._stream_barrier()  # nobody overwrites a shard a peer is still reading
        self.state, self._scratch = self._scratch, self.state

    # ------------------------------------------------------------------ planning
    def _front_end(self, instructions):
        """Host prefix on the pending sparse state + normalisation: ``(ops, state before, state after)``."""
        start = getattr(self, "_pending", None)
        unit_tol = 1e-6 if self.dtype == QB2_C64 else 1e-13
        ops = None
        if start is not None and self.g > 0 and getattr(self, "prefix_support", None) is None and os.environ.get("QB2_PREFIX_SUPPORT") is None:
            # A pass takes QB2_MAX_SPARSE entries PER RANK.  The qubits the prefix spreads out are never dense
            # targets afterwards, so the initial layout puts them on the rank bits and every rank holds 1/world
            # of the list: the host walk may go `world` times further (log2(world) more qubits absorbed -- what
            # keeps GHZ + QFT at two passes however many ranks widen the register).  Checked, not assumed:
            # if some rank's share came out too long, the walk is redone with the single-rank limit.
            stats = dict(self.stats)
            rest = self._take_prefix(list(instructions), prefix.MAX_SUPPORT * self.world)
            ops = normalize_circuit(rest, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
            if self._largest_share(ops) > prefix.MAX_SUPPORT:
                self._pending, ops = start, None
                self.stats.clear()
                self.stats.update(stats)
        if ops is None:
            rest = self._take_prefix(list(instructions))
            ops = normalize_circuit(rest, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
        return ops, start, getattr(self, "_pending", None)

    def compile(self, instructions):
        """Host prefix on the pending sparse state, then plan with swaps (every rank computes the same)."""
        ops, start, entry = self._front_end(instructions)
        return self.compile_ops(ops, start=start, entry=entry)

    def _largest_share(self, ops):
        """Entries of the pending list the fullest rank would hold under the initial layout ``compile_ops``
        is going to choose for ``ops`` (the map is put back)."""
        idx, _ = self._pending
        saved = list(self.pos)
        if os.environ.get("QB2_NO_INITIAL_LAYOUT") is None:
            self._choose_initial_layout(ops)
        phys = prefix.to_physical(idx, self.pos, len(self.pos))
        self.pos[:] = saved
        if not len(phys):
            return 0
        return int(np.bincount((phys >> np.uint64(self.n)).astype(np.int64), minlength=1).max())

    def compile_ops(self, ops, start=None, entry=None):
        """Plan with swaps.  Returns the program; map and state stay what they were (``run`` moves them)."""
        steps, pos_in, pos_before = self._plan_steps(ops)
        prog = ShardedProgram(self, steps, pos_in=pos_in, start=start, entry=entry, pos_before=pos_before)
        self._pending = start
        self.pos[:] = pos_before
        return prog

    def _plan_steps(self, ops):
        """Passes, big steps and swaps for ``ops``: ``(steps, entry map, the caller's map)``; ``self.pos`` is
        left at the exit map."""
        pos_before = list(self.pos)
        if getattr(self, "_pending", None) is not None and self.g > 0 and os.environ.get("QB2_NO_INITIAL_LAYOUT") is None:
            # no data in HBM yet: the map is free.  The program's entry map is the chosen one.
            self._choose_initial_layout(ops)
        pos_in = list(self.pos)
        planner = Planner(self.cfg, self.pos, n_total=self.n_total)
        steps = []
        pending = list(ops)
        protect = set(range(self.cfg.L))
        guard = 0
        while pending:
            part, need = planner.plan(pending)
            steps.extend(part)
            if need is None:
                break
            pending = planner.pending
            guard += 1
            if guard > 10000:
                raise RuntimeError("swap planning does not terminate")
            steps.append(self._plan_swap(pending, protect))
        return steps, pos_in, pos_before

    def _choose_initial_layout(self, ops):
        """The state is still |0...0>, which does not care which qubit sits where: the qubits that are
        needed LAST as dense targets (or never) start on the rank bits, so that the first qubit swap
        happens as late as possible -- for GHZ + QFT one swap instead of two.  Pure bookkeeping:
        ``pos`` entries trade places, no data moves."""
        n = self.n
        first_use = {}
        for i, op in enumerate(ops):
            for q in op.nd:
                first_use.setdefault(q, i)
        nq = len(self.pos)
        order = sorted(range(nq), key=lambda q: (first_use.get(q, len(ops) + 1), q))
        want_global = set(order[nq - self.g :])
        now_global = {q for q in range(nq) if self.pos[q] >= n}
        come_in = sorted(now_global - want_global)
        go_out = sorted(want_global - now_global)
        for a, b in zip(come_in, go_out):
            self.pos[a], self.pos[b] = self.pos[b], self.pos[a]

    def _plan_swap(self, pending, protect):
        n, g = self.n, self.g
        evict = choose_evictions(pending, self.pos, n, g, protect)
        # local permutation that moves evict[i] to position n-g+i and keeps the others in order
        keep = [p for p in range(n) if p not in evict]
        perm = [0] * n  # bit perm[p] of the new index = bit p of the old one
        for newp, oldp in enumerate(keep + evict):
            perm[oldp] = newp
        identity = all(perm[p] == p for p in range(n))
        # relabel: first the local permutation, then top-local <-> rank bits
        newpos = {}
        for q, p in enumerate(self.pos):
            if p < n:
                p2 = perm[p]
                newpos[q] = p2 + g if p2 >= n - g else p2  # evicted positions become rank bits
            else:
                newpos[q] = p - g  # rank bits land on the top local positions
        for q in range(len(self.pos)):
            self.pos[q] = newpos[q]
        return SwapStep(evict, None if identity else perm)

    def apply_circuit(self, instructions):
        """Plan and run in one go, STREAMING like :meth:`StateVector.apply_circuit`: every pass is launched the
        moment the planner has built it, every swap the moment it is decided, so the host plans step k+1 while
        the devices work on step k.  All ranks plan the same, so their collectives line up.  Returns the steps."""
        self._guard()
        ops, _start, _entry = self._front_end(instructions)
        if getattr(self, "_pending", None) is not None and self.g > 0 and os.environ.get("QB2_NO_INITIAL_LAYOUT") is None:
            self._choose_initial_layout(ops)  # nothing in HBM yet: the map is free
        # a pending state is laid out by the map at the entry of the plan (the planner moves ``pos`` on)
        self._entry_pos = list(self.pos) if getattr(self, "_pending", None) is not None else None
        planner = Planner(self.cfg, self.pos, n_total=self.n_total)
        keep = []  # device copies of the blobs stay alive until the launches have been issued

        def launch(step):
            if step.kind == "pass":
                small = int.from_bytes(step.blob[8:12], "little")
                header = ctypes.create_string_buffer(step.blob[:small], small)
                dev = self._upload_blobs([step], [0], len(step.blob))
                keep.append((dev, header))
                self._run_pass(step, header, dev, 0)
                self.stats["passes"] += 1
                self.stats["kernel_ops"] += step.n_ops
                self.stats["gate_ops"] += step.n_gate_ops
                self.stats["phases"] += step.n_phases
            else:
                self._run_big(step)
                self.stats["gate_ops"] += 1
            self._entry_pos = None  # the state is in HBM from here on

        steps = []
        pending = list(ops)
        protect = set(range(self.cfg.L))
        guard = 0
        try:
            while pending:
                part, need = planner.plan(pending, on_step=launch)
                steps.extend(part)
                if need is None:
                    break
                pending = planner.pending
                guard += 1
                if guard > 10000:
                    raise RuntimeError("swap planning does not terminate")
                # (a swap on a state nobody has written yet writes it first, under the entry map: the swap
                # is planned -- and ``pos`` moved -- only after that)
                self._materialize()
                self._entry_pos = None
                swap = self._plan_swap(pending, protect)
                self.run_swap(swap)
                steps.append(swap)
        finally:
            self._entry_pos = None
        self._keep_alive = keep
        return steps

    # ------------------------------------------------------------------ execution of one swap
    def run_swap(self, step):
        if self._peer_setup() and (step.perm is None or self.dtype == QB2_C128 or step.perm[0] == 0):
            self._swap_pull(step.perm)
            self.stats["peer_swaps"] = self.stats.get("peer_swaps", 0) + 1
        else:
            if step.perm is not None:
                self._permute_local(step.perm)
            self._all_to_all()
        self.stats["swaps"] += 1

    # ------------------------------------------------------------------ read-out
    def probabilities(self, qubits):
        self._materialize()
        import torch

        m = len(qubits)
        if m > 30:
            raise MemoryError(f"a dense distribution over {m} qubits does not fit; sample shots instead")
        probs = torch.zeros(1 << m, dtype=torch.float64, device=self.device)
        mes = (ctypes.c_int * max(m, 1))(*[self.pos[q] for q in qubits])
        _lib.check(
            self.lib.qb2_probabilities(self.state.data_ptr(), self.n, self.hi_base, mes, m, probs.data_ptr(), self.dtype, self._stream()),
            "qb2_probabilities",
        )
        self.dist.all_reduce(probs, group=self.group)
        self.io["d2h"] += probs.numel() * 8
        return probs.cpu().numpy()

    def amplitudes(self, indices):
        """Amplitudes at the given basis states (reference order: qubit 0 = most significant bit), on every
        rank: each rank gathers the ones it holds, a sum over ranks completes the list."""
        self._materialize()
        import torch

        idx = np.asarray(indices, dtype=np.uint64)
        nq = self.num_qubits
        phys = np.zeros(idx.shape, dtype=np.uint64)
        for q in range(nq):
            phys |= ((idx >> np.uint64(nq - 1 - q)) & np.uint64(1)) << np.uint64(self.pos[q])
        mine = (phys >> np.uint64(self.n)) == np.uint64(self.rank)
        local = torch.from_numpy((phys & np.uint64((1 << self.n) - 1)).astype(np.int64)).to(self.device)
        vals = self.state.view(-1, 2)[local].to(torch.float64)
        vals = vals * torch.from_numpy(mine.astype(np.float64)).to(self.device)[:, None]
        self.dist.all_reduce(vals, group=self.group)
        v = vals.cpu().numpy()
        return (v[:, 0] + 1j * v[:, 1]).astype(self.np_dtype)

    def norm2(self):
        self._materialize()
        import torch

        out = torch.empty(1, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.qb2_norm2(self.state.data_ptr(), self.n, out.data_ptr(), self.dtype, self._stream()), "qb2_norm2")
        self.dist.all_reduce(out, group=self.group)
        return float(out.item())

    def sample(self, shots, qubits=None, rng=None):
        """Every rank draws its share of the shots from its shard; rank weights come from an
        all-gather of the shard norms.  All ranks must pass identically seeded ``rng``s.
        Returns the outcome integers on every rank (gathered)."""
        self._materialize()
        import torch

        rng = np.random.default_rng(0) if rng is None else rng
        if qubits is None:
            qubits = list(range(self.num_qubits))
        shots = int(shots)
        cb = self.lib.qb2_sample_chunk_bits(self.n)
        sums = torch.empty(1 << (self.n - cb), dtype=torch.float64, device=self.device)
        _lib.check(
            self.lib.qb2_sample_prepare(self.state.data_ptr(), self.n, sums.data_ptr(), self.dtype, self._stream()),
            "qb2_sample_prepare",
        )
        weights = torch.zeros(self.world, dtype=torch.float64, device=self.device)
        weights[self.rank] = sums[-1]
        self.dist.all_reduce(weights, group=self.group)
        w = weights.cpu().numpy()
        split = rng.multinomial(shots, w / w.sum())  # same on every rank (same seed)
        uniforms = rng.random(shots)
        lo = int(split[: self.rank].sum())
        mine = np.ascontiguousarray(uniforms[lo : lo + int(split[self.rank])])
        phys_all = torch.zeros(shots, dtype=torch.int64, device=self.device)
        if len(mine):
            u_dev = torch.from_numpy(mine).to(self.device)
            _lib.check(
                self.lib.qb2_sample_draw(
                    self.state.data_ptr(), self.n, sums.data_ptr(), u_dev.data_ptr(), len(mine),
                    phys_all[lo : lo + len(mine)].data_ptr(), self.dtype, self._stream(),
                ),
                "qb2_sample_draw",
            )
            phys_all[lo : lo + len(mine)] |= self.hi_base
        self.dist.all_reduce(phys_all, group=self.group)  # disjoint slices: a sum is a gather
        self.io["h2d"] += mine.nbytes
        self.io["d2h"] += shots * 8
        phys = phys_all.cpu().numpy().astype(np.uint64)
        m = len(qubits)
        out = np.zeros(phys.shape, dtype=np.uint64)
        for k, q in enumerate(qubits):
            out |= ((phys >> np.uint64(self.pos[q])) & np.uint64(1)) << np.uint64(m - 1 - k)
        return out

    def statevector(self):
        """Gather the whole state in canonical order on every rank (small registers only)."""
        self._materialize()
        import torch

        if self.n_total > 30:
            raise MemoryError("statevector(): the gathered state would not fit on one device")
        shards = [torch.empty_like(self.state) for _ in range(self.world)]
        self.dist.all_gather(shards, self.state, group=self.group)
        full = torch.cat(shards).cpu().numpy().view(self.np_dtype)
        return _canonical(full, self.pos, self.num_qubits)


class PlanOnlyShard(ShardedStateVector):
    """The planning half of a sharded state, without a device or a process group: what each of ``world`` ranks
    plans for a circuit (all ranks plan the same).  For compiling specialised kernels ahead of time
    (:func:`qrisp_b200.jit.prebuild`) and for looking at plans (``tools/dump_plan.py``); it cannot run anything."""

    def __init__(self, num_qubits, world=1, dtype=np.complex64, T=None, r=None, L=4, tma=None, prefix_support=None):
        self._init_host(num_qubits, world, 0, dtype, T, r, L, tma, prefix_support)
        self.specialize = False
        self._pending = prefix.zero_state()

    def _setup_device(self, device):
        raise _lib.Qb2Error("PlanOnlyShard plans; it has no device")

    def plan(self, instructions):
        """``(steps, entry state, entry map)`` for unitary ``instructions`` applied to |0...0>."""
        self.reset()
        ops, start, entry = self._front_end(instructions)
        steps, pos_in, pos_before = self._plan_steps(ops)
        self._pending = start
        self.pos[:] = pos_before
        return steps, entry, pos_in


def _canonical(full, pos, nq):
    N = len(pos)
    j = np.arange(1 << N, dtype=np.int64)
    i = np.zeros_like(j)
    for q in range(N):
        i |= ((j >> pos[q]) & 1) << (nq - 1 - q if q < nq else q)
    out = np.empty_like(full)
    out[i] = full
    return out[: 1 << nq].copy()


class ShardedProgram(Program):
    """Passes + swaps.  Pass blobs are uploaded once; swaps run on the same stream."""

    def __init__(self, sv, steps, pos_in=None, start=None, entry=None, pos_before=None):
        self._swap_steps = [s for s in steps if s.kind == "swap"]
        super().__init__(sv, steps, pos_in=pos_in, start=start, entry=entry)
        # the map the caller has when the program starts; with a pending state the program is free to
        # begin from another one (pos_in, see ShardedStateVector._choose_initial_layout)
        self.pos_before = list(self.pos_in) if pos_before is None else list(pos_before)
        self.n_swaps = len(self._swap_steps)
        self.comm_ms = None

    def _enter(self):
        sv = self.sv
        if list(sv.pos) != self.pos_before:
            raise _lib.Qb2Error("this program was compiled against another qubit map (compile it where it runs, in order)")
        if self.pos_before != self.pos_in:
            if getattr(sv, "_pending", None) is None:
                raise _lib.Qb2Error("this program chooses its own entry layout and needs a state that has not been written yet")
            sv.pos[:] = self.pos_in
        super()._enter()

    def run(self, time_comm=False):
        sv = self.sv
        self._enter()
        comm = []
        first = True
        for s, o, h in zip(self.steps, self.offsets, self._headers):
            if s.kind == "pass":
                sv._run_pass(s, h, self.dev, o, sparse=self._entry_sparse if first else None)
                first = False
            elif s.kind == "swap":
                first = False
                if time_comm:
                    import torch

                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    sv.run_swap(s)
                    e1.record()
                    comm.append((e0, e1))
                else:
                    sv.run_swap(s)
            else:
                sv._run_big(s)
        if time_comm:
            import torch

            torch.cuda.synchronize()
            self.comm_ms = sum(a.elapsed_time(b) for a, b in comm)
        sv.pos[:] = self.pos_out
        sv.stats["passes"] += self.n_passes
        sv.stats["kernel_ops"] += self.n_kernel_ops
        sv.stats["gate_ops"] += self.n_gate_ops
        sv.stats["phases"] += self.n_phases


def sample_counts_sharded(qc, shots, seed=0, group=None, dtype=np.complex64, **engine_kwargs):
    """Multi-GPU counterpart of :func:`qrisp_b200.simulator.sample_counts`: every rank calls
    it with the same arguments; the counts dict is returned on every rank."""
    from .circuit import Circuit
    from .preprocess import unitarize

    qc = qc if isinstance(qc, Circuit) else Circuit.from_qrisp(qc)
    instrs, n, measurements = unitarize(qc)
    sv = ShardedStateVector(n, group=group, dtype=dtype, **engine_kwargs)
    sv.apply_circuit(instrs)
    measurements.sort(key=lambda x: -x[1])
    mes_qubits = [q for q, _ in measurements]
    outs = sv.sample(int(shots), mes_qubits, rng=np.random.default_rng(seed))
    from .simulator import counts_dict

    return counts_dict(outs, len(mes_qubits)), sv
