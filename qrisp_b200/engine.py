"""Device-side statevector: owns the HBM buffers (as torch tensors) and drives libqb2.so.

This is the B200 counterpart of the reference's ``QuantumState`` / ``TensorFactor`` /
``BiArray`` stack (reference: quantum_state.py:34-245, tensor_factor.py:28-273,
bi_arrays.py:702-1115): one dense array of ``2**n`` amplitudes in HBM, swept by fused
multi-gate passes.  PyTorch is used for buffer ownership, streams and (multi-GPU)
``torch.distributed``; every amplitude is touched by hand-written CUDA only.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib, prefix
from .normalize import normalize_circuit
from .planner import QB2_C64, QB2_C128, PlanConfig, Planner

QB2_FEATURE_ZERO_INPUT = 4  # include/qrisp_b200.h

MIN_QUBITS = 9  # smaller registers are padded with idle |0> qubits (a 4 KB state)


def _torch():
    import torch

    return torch


def _specialize_mode(arg):
    if arg is not None:
        return bool(arg)
    env = os.environ.get("QB2_SPECIALIZE")
    if env is None or env == "":
        return None
    return env != "0"


class DeviceOps:
    """How a state object talks to libqb2.so; shared by the single-device and the sharded
    state.  (CPU tests of the sharding logic substitute these four methods.)"""

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _upload_blobs(self, steps, offs, total):
        torch = _torch()
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        for s, o in zip(steps, offs):
            if o is not None:
                hv[o : o + len(s.blob)] = np.frombuffer(s.blob, dtype=np.uint8)
        dev = host.to(self.device, non_blocking=True)
        # keep the staging buffers alive until the copies have run (bounded: the oldest go first)
        pinned = getattr(self, "_pinned", None)
        if not isinstance(pinned, list):
            pinned = self._pinned = []
        pinned.append(host)
        del pinned[:-64]
        self.io["h2d"] += total
        return dev

    def _guard(self):
        """Launches go to the CURRENT device: make it this state's (two engines on two devices may
        alternate in one process)."""
        if self.device.type == "cuda":
            _torch().cuda.set_device(self.device)

    def _upload_bytes(self, data):
        torch = _torch()
        host = torch.empty(max(len(data), 16), dtype=torch.uint8, pin_memory=True)
        host.numpy()[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        dev = host.to(self.device, non_blocking=True)
        pinned = getattr(self, "_pinned", None)
        if not isinstance(pinned, list):
            pinned = self._pinned = []
        pinned.append(host)
        del pinned[:-64]
        self.io["h2d"] += len(data)
        return dev

    def _pending_map(self):
        """The qubit map a pending state is laid out by: the map at the ENTRY of the plan that is being
        launched (the planner has already moved ``self.pos`` on by the time a step is launched)."""
        m = getattr(self, "_entry_pos", None)
        return self.pos if m is None else m

    def _sparse_for(self, step, state):
        """Device copy of the sparse input list of ``step`` for the pending state ``state``."""
        idx, amp = state
        phys = prefix.to_physical(idx, self._pending_map(), len(self.pos))
        data, count = prefix.sparse_input_list(step, phys, amp, self.n, self.hi_base, self.np_dtype)
        return (self._upload_bytes(data) if count else None), count

    def _run_pass(self, step, header, dev, offset, sparse=None, norms=False):
        """``sparse``: a ready ``(device list, count)`` for a pass that starts from the pending state
        (a replayed program keeps it resident); built here otherwise.  ``norms``: let the last pass of a plan
        leave the sampler's per-tile sums (the streaming path of ``run()`` asks for them; a replayed program
        does when its ``want_norms`` is set)."""
        sp_ptr, sp_count = None, 0
        pending = getattr(self, "_pending", None)
        if pending is not None:
            # the state buffer has not been written since the reset: the pass builds its tiles from the
            # list of non-zero amplitudes (QB2_FEATURE_ZERO_INPUT, header.features at byte 56) instead of
            # reading 2**n zeros, and stores all-zero tiles without running the ops
            if sparse is None:
                sparse = self._sparse_for(step, pending)
            if sparse[1] > prefix.MAX_SUPPORT:
                # more entries on this rank than a pass takes (a sharded list is `world` times longer and an
                # earlier call may have left it pending under another layout): write the state, run normally
                self._materialize()
                return self._run_pass(step, header, dev, offset, norms=norms)
            self._sparse_keep = sparse
            sp_ptr = sparse[0].data_ptr() if sparse[0] is not None else None
            sp_count = sparse[1]
            raw = bytearray(header.raw)
            raw[56:60] = (int.from_bytes(raw[56:60], "little") | QB2_FEATURE_ZERO_INPUT).to_bytes(4, "little")
            header = ctypes.create_string_buffer(bytes(raw), len(raw))
            self._pending = None
        kernel = getattr(step, "jit", None)
        mode = getattr(self, "specialize", False)
        if kernel is None and mode is not False and not hasattr(step, "jit"):
            from . import jit

            # True: compiled once per pass structure (seconds), then cached; None (the default): only what the
            # cache already holds -- a kernel somebody has paid for is used, nothing is ever compiled unasked
            kernel = step.jit = jit.kernel_for(step, build=mode is True)
        # the last pass of a plan, when its tile is the lowest T positions, also leaves the per-tile |amp|**2 sums:
        # the sampler's first level, for free (see sample())
        self._norms = None
        want, norms = norms, None
        T = len(step.tile_pos)
        if (want and getattr(step, "last", False) and getattr(self, "fuse_norms", True) and getattr(self, "world", 1) == 1 and self.n > T
                and list(step.tile_pos) == list(range(T))):
            norms = _torch().zeros(1 << (self.n - T), dtype=_torch().float64, device=self.device)
        _lib.check(
            self.lib.qb2_run_pass_ex(
                self._raw_state().data_ptr(), self.n, self.hi_base, header, dev.data_ptr() + offset, self.dtype, self._stream(),
                sp_ptr, sp_count, kernel, norms.data_ptr() if norms is not None else None,
            ),
            "qb2_run_pass",
        )
        if kernel:
            self.stats["jit_passes"] = self.stats.get("jit_passes", 0) + 1
        if norms is not None:
            self._norms = (norms, T)

    def _raw_state(self):
        """The amplitude buffer as it is (a pass may run on a state whose zeros were never written)."""
        return self.state

    def _materialize(self):
        """Write the pending sparse state for real (see :meth:`StateVector.reset`)."""
        pending = getattr(self, "_pending", None)
        if pending is not None:
            self._norms = None
            self._pending = None
            self._init_sparse(pending)

    def _init_sparse(self, state):
        torch = _torch()
        idx, amp = state
        phys = prefix.to_physical(idx, self._pending_map(), len(self.pos))
        if len(phys):
            mine = (phys >> np.uint64(self.n)) == np.uint64(self.hi_base >> self.n)
            phys, amp = phys[mine] & np.uint64((1 << self.n) - 1), amp[mine]
        count = len(phys)
        buf = self._raw_state()
        if count:
            data = phys.astype(np.uint64).tobytes()
            pad = (len(data) + 15) // 16 * 16
            data = data + b"\0" * (pad - len(data)) + amp.astype(self.np_dtype).tobytes()
            dev = self._upload_bytes(data)
            self._sparse_keep = (dev, count)
            ip, ap = dev.data_ptr(), dev.data_ptr() + pad
        else:
            ip = ap = None
        _lib.check(self.lib.qb2_init_sparse(buf.data_ptr(), self.n, ip, ap, count, self.dtype, self._stream()), "qb2_init_sparse")

    def _take_prefix(self, instructions, support=None):
        """Evolve the pending sparse state through a prefix of ``instructions`` on the host
        (:mod:`qrisp_b200.prefix`); returns the instructions that are left for the dense sweeps.
        ``support``: entries the list may grow to (default: this engine's ``prefix_support``)."""
        pending = getattr(self, "_pending", None)
        if pending is None or not instructions:
            return instructions
        if support is None:
            support = getattr(self, "prefix_support", None)
        consumed, state = prefix.sparse_prefix(instructions, pending, support)
        if consumed:
            self._pending = state
            self.stats["prefix_gates"] = self.stats.get("prefix_gates", 0) + consumed
            self.stats["gate_ops"] += consumed
        return instructions[consumed:]

    def _prepare_block_tc(self, s):
        """Host side of a 5-qubit tensor-core block: targets sorted by position, the matrix re-indexed to match,
        split into its two TF32 halves in the kernel's layout (qb2_pack_block_tc) and uploaded.  Kept on the
        step: a replayed program pays it once."""
        ready = getattr(s, "_tc_ready", None)
        if ready is not None and ready[0] is self:
            return ready[1:]
        order = np.argsort(s.positions)
        positions = [int(s.positions[i]) for i in order]
        old = np.arange(32)
        new = np.zeros(32, dtype=np.int64)  # matrix index after sorting the targets by position
        for j in range(5):
            new |= ((old >> int(order[j])) & 1) << j
        m = np.empty((32, 32), dtype=np.complex128)
        m[np.ix_(new, new)] = np.asarray(s.matrix, dtype=np.complex128)
        packed = np.empty(_lib.QB2_BLOCK_TC_BYTES // 4, dtype=np.float32)
        _lib.check(self.lib.qb2_pack_block_tc(np.ascontiguousarray(m).ctypes.data, packed.ctypes.data), "qb2_pack_block_tc")
        dev = self._upload_bytes(packed.tobytes())
        tg = (ctypes.c_int * 5)(*positions)
        try:
            s._tc_ready = (self, dev, tg)
        except AttributeError:
            pass
        return dev, tg

    def _run_block_tc(self, s):
        """A dense 5-qubit block on the tensor cores (qb2_apply_block_tc: tcgen05.mma, TF32 x 3), in place."""
        self._norms = None
        dev, tg = self._prepare_block_tc(s)
        _lib.check(
            self.lib.qb2_apply_block_tc(self._raw_state().data_ptr(), self.n, self.hi_base, dev.data_ptr(), tg, s.cmask, s.cval,
                                        self._stream()),
            "qb2_apply_block_tc",
        )
        self.stats["big"] += 1
        self.stats["block_tc"] = self.stats.get("block_tc", 0) + 1

    def _csd_steps(self, matrix, positions, cmask, cval, out):
        """Cosine-sine decomposition down to 5-qubit blocks.  ``matrix``: index bit i <-> positions[i].  Appends
        ("tc", positions, matrix, cmask, cval) and ("mux", target, controls, theta, cmask, cval) in the order they
        act on the state:  U = diag(u1, u2) . CS(theta) . diag(v1h, v2h), split on the top index bit."""
        k = len(positions)
        if k == 5:
            out.append(("tc", list(positions), matrix, cmask, cval))
            return
        from scipy.linalg import cossin

        half = 1 << (k - 1)
        (u1, u2), theta, (v1h, v2h) = cossin(np.asarray(matrix, dtype=np.complex128), p=half, q=half, separate=True)
        top, rest = int(positions[k - 1]), list(positions[: k - 1])
        bit = 1 << top
        self._csd_steps(v1h, rest, cmask | bit, cval, out)
        self._csd_steps(v2h, rest, cmask | bit, cval | bit, out)
        out.append(("mux", top, rest, np.asarray(theta, dtype=np.float64), cmask, cval))
        self._csd_steps(u1, rest, cmask | bit, cval, out)
        self._csd_steps(u2, rest, cmask | bit, cval | bit, out)

    def _run_block_csd(self, s):
        """A dense block on 6..8 qubits: controlled 5-qubit blocks on the tensor cores around multiplexed real
        rotations (qb2_apply_mux_ry), from the cosine-sine decomposition of the matrix (host, scipy).  The
        decomposition and its device operands are kept on the step: a replayed program pays them once."""
        from .planner import BigMatStep

        ready = getattr(s, "_csd_ready", None)
        if ready is None or ready[0] is not self:
            steps = []
            self._csd_steps(np.asarray(s.matrix, dtype=np.complex128), [int(p) for p in s.positions], int(s.cmask), int(s.cval), steps)
            real = np.float32 if self.dtype == QB2_C64 else np.float64
            prepared = []
            for st in steps:
                if st[0] == "tc":
                    prepared.append(("tc", BigMatStep(st[1], st[2], st[3], st[4])))
                else:
                    _kind, target, ctrl, theta, cmask, cval = st
                    cs = np.stack([np.cos(theta), np.sin(theta)], axis=1).astype(real)
                    prepared.append(("mux", target, (ctypes.c_int * len(ctrl))(*ctrl), len(ctrl), self._upload_bytes(cs.tobytes()), cmask, cval))
            ready = (self, prepared)
            try:
                s._csd_ready = ready
            except AttributeError:
                pass
        for st in ready[1]:
            if st[0] == "tc":
                self._run_block_tc(st[1])
                self.stats["big"] -= 1
            else:
                _kind, target, arr, nc, dev, cmask, cval = st
                _lib.check(
                    self.lib.qb2_apply_mux_ry(self._raw_state().data_ptr(), self.n, self.hi_base, target, arr, nc, dev.data_ptr(),
                                              cmask, cval, self.dtype, self._stream()),
                    "qb2_apply_mux_ry",
                )
        self._norms = None
        self.stats["big"] += 1
        self.stats["block_csd"] = self.stats.get("block_csd", 0) + 1

    def _run_big(self, s):
        self._materialize()
        self._norms = None
        torch = _torch()
        k = len(s.positions)
        if (6 <= k <= 8 and self.dtype == QB2_C64 and self.n >= 12 and all(p < self.n for p in s.positions)
                and getattr(self, "block_tc", True) and os.environ.get("QB2_NO_BLOCK_TC") is None):
            return self._run_block_csd(s)
        if (k == 5 and self.dtype == QB2_C64 and self.n >= 12 and all(p < self.n for p in s.positions)
                and getattr(self, "block_tc", True) and os.environ.get("QB2_NO_BLOCK_TC") is None):
            return self._run_block_tc(s)
        m = np.ascontiguousarray(s.matrix, dtype=self.np_dtype)
        m_dev = self._upload_bytes(m.tobytes())
        tg = (ctypes.c_int * k)(*s.positions)
        dst = self._scratch_buf()
        src = self._raw_state()
        _lib.check(
            self.lib.qb2_apply_matrix_big(
                src.data_ptr(), dst.data_ptr(), self.n, self.hi_base, m_dev.data_ptr(), tg, k, s.cmask, s.cval,
                self.dtype, self._stream(),
            ),
            "qb2_apply_matrix_big",
        )
        self.state, self._scratch = dst, src
        self.stats["big"] += 1


class StateVector(DeviceOps):
    """``num_qubits`` logical qubits in |0...0>.

    ``pos[q]`` is the physical position (bit of the array index) of logical qubit ``q``;
    the reference's canonical order is ``pos[q] = num_qubits - 1 - q``.
    """

    def __init__(self, num_qubits, dtype=np.complex64, device=None, T=None, r=None, L=4, tma=None, prefix_support=None,
                 specialize=None):
        torch = _torch()
        # specialised pass kernels (qrisp_b200.jit).  True: compile what is missing (seconds per pass, once);
        # False: interpreter only; None (default): use kernels that are already in the cache, compile nothing
        self.specialize = _specialize_mode(specialize)
        if not torch.cuda.is_available():
            raise _lib.Qb2Error("the B200 engine needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.num_qubits = int(num_qubits)
        self.n = max(self.num_qubits, MIN_QUBITS)
        dtype = np.dtype(dtype)
        if dtype == np.complex64:
            self.dtype, self.np_dtype, real = QB2_C64, np.complex64, torch.float32
        elif dtype == np.complex128:
            self.dtype, self.np_dtype, real = QB2_C128, np.complex128, torch.float64
        else:
            raise TypeError(f"unsupported amplitude dtype {dtype}")
        self.real = real
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        self.cfg = PlanConfig(self.n, self.dtype, T=T, r=r, L=L, tma=tma)
        self.prefix_support = prefix_support  # None: the default of qrisp_b200.prefix; 0: no host prefix
        self.pos = [self.num_qubits - 1 - q for q in range(self.num_qubits)] + list(range(self.num_qubits, self.n))
        self.hi_base = 0
        self._buf = torch.empty(2 << self.n, dtype=real, device=self.device)
        self._scratch = None
        self.stats = {"passes": 0, "kernel_ops": 0, "gate_ops": 0, "phases": 0, "big": 0}
        self.io = {"h2d": 0, "d2h": 0}  # bytes copied between host and device
        self.reset()

    # ------------------------------------------------------------------ plumbing
    @property
    def state(self):
        """The amplitude buffer (torch tensor of 2 * 2**n reals).  Looking at it writes a pending
        |0...0> first (see :meth:`reset`)."""
        self._materialize()
        return self._buf

    @state.setter
    def state(self, buf):
        self._buf = buf

    def _raw_state(self):
        return self._buf

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _scratch_buf(self):
        if self._scratch is None:
            self._scratch = _torch().empty_like(self._buf)
        return self._scratch

    def reset(self):
        """Back to |0...0>.  Lazy: the state is PENDING -- a short list of non-zero amplitudes on the
        host (:mod:`qrisp_b200.prefix`).  Gates are applied to the list while it stays short; the first
        dense pass builds its tiles from it instead of reading 2**n zeros; anything else that looks at
        the buffer first makes :meth:`_materialize` write it."""
        self._pending = prefix.zero_state()
        self._norms = None
        # no data sits in HBM any more, so the qubit map is free again: back to the canonical one (a
        # reset engine is a new engine; compiled programs name the map they start from)
        nq = self.num_qubits
        self.pos[:] = [nq - 1 - q for q in range(nq)] + list(range(nq, self.n))

    def synchronize(self):
        _torch().cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ gates
    def compile(self, instructions):
        """Normalise + plan a list of unitary instructions against the CURRENT qubit map (and, when the
        state is still pending, the current sparse state: the program then starts with the host
        prefix).  Returns a compiled program that :meth:`execute` runs; map and state stay what they were
        (running the program moves them)."""
        pos_in = list(self.pos)
        start = self._pending
        instructions = self._take_prefix(list(instructions))
        entry = self._pending
        # complex128 keeps non-unit diagonal entries exactly as given (see normalize.py)
        unit_tol = 1e-6 if self.dtype == QB2_C64 else 1e-13
        ops = normalize_circuit(instructions, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
        prog = self.compile_ops(ops, pos_in=pos_in, start=start, entry=entry)
        return prog

    def compile_ops(self, ops, pos_in=None, start=None, entry=None):
        pos_in = list(self.pos) if pos_in is None else pos_in
        planner = Planner(self.cfg, self.pos)
        steps, need = planner.plan(ops)
        if need is not None:
            raise _lib.Qb2Error(f"op needs non-local positions {need} on a single-device state")
        prog = Program(self, steps, pos_in=pos_in, start=start, entry=entry)
        self._pending = start  # nothing has run yet: the caller's state is what it was
        self.pos[:] = pos_in
        return prog

    def execute(self, program):
        program.run()

    def apply_circuit(self, instructions):
        """Plan and run in one go, STREAMING: every pass is launched the moment the planner has built
        it (launches are asynchronous), so the host plans pass k+1 while the device sweeps pass k and
        planning disappears from the wall time except for the first pass.  Returns the step list."""
        self._guard()
        instructions = self._take_prefix(list(instructions))
        unit_tol = 1e-6 if self.dtype == QB2_C64 else 1e-13
        ops = normalize_circuit(instructions, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
        planner = Planner(self.cfg, self.pos)
        keep = []  # device copies of the blobs stay alive until the launches have been issued
        self._entry_pos = list(self.pos) if self._pending is not None else None

        def launch(step):
            if step.kind == "pass":
                small = int.from_bytes(step.blob[8:12], "little")
                header = ctypes.create_string_buffer(step.blob[:small], small)
                dev = self._upload_blobs([step], [0], len(step.blob))
                keep.append((dev, header))
                self._run_pass(step, header, dev, 0, norms=True)
                self.stats["passes"] += 1
                self.stats["kernel_ops"] += step.n_ops
                self.stats["gate_ops"] += step.n_gate_ops
                self.stats["phases"] += step.n_phases
            else:
                self._run_big(step)
                self.stats["gate_ops"] += 1

        try:
            steps, need = planner.plan(ops, on_step=launch)
        finally:
            self._entry_pos = None
        if need is not None:
            raise _lib.Qb2Error(f"op needs non-local positions {need} on a single-device state")
        self._keep_alive = keep
        return steps

    def apply_matrix(self, matrix, qubits):
        """Drop-in for ``TensorFactor.apply_matrix`` (tensor_factor.py:67-89): ``matrix`` is
        big endian in ``qubits``."""
        from .circuit import Instruction

        return self.apply_circuit([Instruction("unitary", qubits, (), matrix)])

    # ------------------------------------------------------------------ read-out
    def statevector(self):
        """The amplitudes in the reference's canonical order (qubit 0 = most significant
        bit; simulator.py:288 / tensor_factor.py:163-169), as a host numpy array."""
        self._materialize()
        torch = _torch()
        nq = self.num_qubits
        perm = (ctypes.c_int * self.n)()
        for q in range(self.n):
            perm[self.pos[q]] = nq - 1 - q if q < nq else q
        if all(perm[p] == p for p in range(self.n)):
            src = self.state
        else:
            src = self._scratch_buf()
            _lib.check(
                self.lib.qb2_permute_bits(self.state.data_ptr(), src.data_ptr(), self.n, perm, self.dtype, self._stream()),
                "qb2_permute_bits",
            )
        host = src[: 2 << nq].cpu().numpy()
        self.io["d2h"] += host.nbytes
        return host.view(self.np_dtype).copy()

    def amplitudes(self, indices):
        """Amplitudes at the given basis states (integers in the reference's order: qubit 0 = most
        significant bit) -- a gather of a few entries, for spot checks of states too large to copy out."""
        self._materialize()
        torch = _torch()
        idx = np.asarray(indices, dtype=np.uint64)
        nq = self.num_qubits
        phys = np.zeros(idx.shape, dtype=np.uint64)
        for q in range(nq):
            phys |= ((idx >> np.uint64(nq - 1 - q)) & np.uint64(1)) << np.uint64(self.pos[q])
        sel = torch.from_numpy(phys.astype(np.int64)).to(self.device)
        vals = self.state.view(-1, 2)[sel].cpu().numpy()
        self.io["h2d"] += phys.nbytes
        self.io["d2h"] += vals.nbytes
        return (vals[:, 0] + 1j * vals[:, 1]).astype(self.np_dtype)

    def norm2(self):
        self._materialize()
        torch = _torch()
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.qb2_norm2(self.state.data_ptr(), self.n, out.data_ptr(), self.dtype, self._stream()), "qb2_norm2")
        return float(out.item())

    def probabilities(self, qubits, to_host=True):
        """Marginal distribution of ``qubits``; ``qubits[0]`` is the most significant bit of
        the outcome index (reference: tensor_factor.py:171-182)."""
        self._materialize()
        torch = _torch()
        m = len(qubits)
        if m > 30:
            raise MemoryError(f"a dense distribution over {m} qubits does not fit; sample shots instead")
        probs = torch.zeros(1 << m, dtype=torch.float64, device=self.device)
        mes = (ctypes.c_int * max(m, 1))(*[self.pos[q] for q in qubits])
        _lib.check(
            self.lib.qb2_probabilities(
                self.state.data_ptr(), self.n, self.hi_base, mes, m, probs.data_ptr(), self.dtype, self._stream()
            ),
            "qb2_probabilities",
        )
        if not to_host:
            return probs
        self.io["d2h"] += probs.numel() * 8
        return probs.cpu().numpy()

    def sample(self, shots, qubits=None, rng=None, uniforms=None):
        """Draw ``shots`` basis states from |amp|^2 and return the outcome integers of
        ``qubits`` (``qubits[0]`` most significant).  Host RNG, so a seed fixes the result."""
        self._materialize()
        torch = _torch()
        if uniforms is None:
            rng = np.random.default_rng() if rng is None else rng
            uniforms = rng.random(int(shots))
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        shots = len(uniforms)
        if qubits is None:
            qubits = list(range(self.num_qubits))
        fused = getattr(self, "_norms", None)
        if fused is not None:
            # the last pass left the per-tile |amp|**2 sums: no read of the whole state (a copy is scanned, the sums
            # stay valid for further calls)
            cb = fused[1]
            sums = fused[0].clone()
            _lib.check(self.lib.qb2_sample_scan(sums.data_ptr(), sums.numel(), self._stream()), "qb2_sample_scan")
            self.stats["fused_norms"] = self.stats.get("fused_norms", 0) + 1
        else:
            cb = self.lib.qb2_sample_chunk_bits(self.n)
            sums = torch.empty(1 << (self.n - cb), dtype=torch.float64, device=self.device)
            _lib.check(
                self.lib.qb2_sample_prepare(self.state.data_ptr(), self.n, sums.data_ptr(), self.dtype, self._stream()),
                "qb2_sample_prepare",
            )
        u_dev = torch.from_numpy(uniforms).to(self.device)
        idx = torch.empty(shots, dtype=torch.int64, device=self.device)
        _lib.check(
            self.lib.qb2_sample_draw_chunks(
                self.state.data_ptr(), self.n, cb, sums.data_ptr(), u_dev.data_ptr(), shots, idx.data_ptr(), self.dtype,
                self._stream(),
            ),
            "qb2_sample_draw_chunks",
        )
        self.io["h2d"] += uniforms.nbytes
        self.io["d2h"] += shots * 8
        phys = idx.cpu().numpy().astype(np.uint64) | np.uint64(self.hi_base)
        return self.outcomes_from_physical(phys, qubits)

    def outcomes_from_physical(self, phys, qubits):
        m = len(qubits)
        out = np.zeros(phys.shape, dtype=np.uint64)
        for k, q in enumerate(qubits):
            out |= ((phys >> np.uint64(self.pos[q])) & np.uint64(1)) << np.uint64(m - 1 - k)
        return out


class Program:
    """A planned list of steps with its pass blobs resident on the device.

    A program is compiled against a qubit map (``pos_in``) and, when the state was still pending,
    against that sparse state (``start``): :meth:`run` checks both, because the passes bake in physical
    positions -- run against another map they would silently hit the wrong qubits.  ``entry`` is the
    sparse state after the host prefix, which the first pass starts from; its device list is built once.
    """

    def __init__(self, sv, steps, pos_in=None, start=None, entry=None):
        self.sv = sv
        self.steps = steps
        self.pos_in = list(sv.pos) if pos_in is None else list(pos_in)
        self.pos_out = list(sv.pos)
        self.start, self.entry = start, entry
        offs, total = [], 0
        for s in steps:
            if s.kind == "pass":
                offs.append(total)
                total += len(s.blob)
            else:
                offs.append(None)
        self.offsets = offs
        self.h2d_bytes = total
        self.dev = sv._upload_blobs(steps, offs, total) if total else None
        self._headers = []
        for s in steps:
            # the small section (header.small_bytes) is handed to the launch by value
            if s.kind == "pass":
                small = int.from_bytes(s.blob[8:12], "little")
                self._headers.append(ctypes.create_string_buffer(s.blob[:small], small))
            else:
                self._headers.append(None)
        self.want_norms = False  # True: the last pass also leaves the sampler's per-tile sums (see StateVector.sample)
        self._entry_sparse = None
        if entry is not None and steps and steps[0].kind == "pass":
            sv._entry_pos = self.pos_in  # the first pass sees the entry map (relabels come later)
            self._entry_sparse = sv._sparse_for(steps[0], entry)
            sv._entry_pos = None
        self.n_passes = sum(1 for s in steps if s.kind == "pass")
        self.n_kernel_ops = sum(s.n_ops for s in steps if s.kind == "pass")
        self.n_gate_ops = sum(s.n_gate_ops for s in steps if s.kind == "pass") + sum(1 for s in steps if s.kind == "big")
        self.n_phases = sum(s.n_phases for s in steps if s.kind == "pass")

    def specialize(self):
        """Compile (or fetch from the cache) a specialised kernel for every TMA pass of this program
        (:mod:`qrisp_b200.jit`); returns how many passes have one.  Replays then run without the interpreter."""
        from . import jit

        passes = [s for s in self.steps if s.kind == "pass"]
        for s, k in zip(passes, jit.kernels_for(passes)):
            s.jit = k
        return sum(1 for s in passes if s.jit)

    def _enter(self):
        sv = self.sv
        sv._guard()
        if list(sv.pos) != self.pos_in:
            raise _lib.Qb2Error("this program was compiled against another qubit map (compile it where it runs, in order)")
        if self.start is not None:
            now = getattr(sv, "_pending", None)
            if now is None or not (np.array_equal(now[0], self.start[0]) and np.array_equal(now[1], self.start[1])):
                raise _lib.Qb2Error("this program was compiled for a state that has not been written yet (reset() first)")
            sv._pending = self.entry
        elif getattr(sv, "_pending", None) is not None:
            sv._materialize()

    def run(self):
        sv = self.sv
        self._enter()
        first = True
        for s, o, h in zip(self.steps, self.offsets, self._headers):
            if s.kind == "pass":
                sv._run_pass(s, h, self.dev, o, sparse=self._entry_sparse if first else None, norms=self.want_norms)
            else:
                sv._run_big(s)
            first = False
        sv.pos[:] = self.pos_out  # (a state that is still pending is laid out by whatever map is current)
        sv.stats["passes"] += self.n_passes
        sv.stats["kernel_ops"] += self.n_kernel_ops
        sv.stats["gate_ops"] += self.n_gate_ops
        sv.stats["phases"] += self.n_phases
