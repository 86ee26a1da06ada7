"""Device-side statevector: owns the HBM buffers (as torch tensors) and drives libqb2.so.

This is the B200 counterpart of the reference's ``QuantumState`` / ``TensorFactor`` /
``BiArray`` stack (reference: quantum_state.py:34-245, tensor_factor.py:28-273,
bi_arrays.py:702-1115): one dense array of ``2**n`` amplitudes in HBM, swept by fused
multi-gate passes.  PyTorch is used for buffer ownership, streams and (multi-GPU)
``torch.distributed``; every amplitude is touched by hand-written CUDA only.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .normalize import normalize_circuit
from .planner import QB2_C64, QB2_C128, PlanConfig, Planner

QB2_FEATURE_ZERO_INPUT = 4  # include/qrisp_b200.h

MIN_QUBITS = 9  # smaller registers are padded with idle |0> qubits (a 4 KB state)


def _torch():
    import torch

    return torch


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

    def _run_pass(self, step, header, dev, offset):
        if getattr(self, "_fresh", False):
            # the state is |0...0> and not written yet: the pass synthesises its input
            # (QB2_FEATURE_ZERO_INPUT, header.features at byte 56) instead of reading 2**n zeros
            raw = bytearray(header.raw)
            raw[56:60] = (int.from_bytes(raw[56:60], "little") | QB2_FEATURE_ZERO_INPUT).to_bytes(4, "little")
            header = ctypes.create_string_buffer(bytes(raw), len(raw))
            self._fresh = False
        _lib.check(
            self.lib.qb2_run_pass(
                self._raw_state().data_ptr(), self.n, self.hi_base, header, dev.data_ptr() + offset, self.dtype, self._stream()
            ),
            "qb2_run_pass",
        )

    def _raw_state(self):
        """The amplitude buffer as it is (a pass may run on a fresh state without the zeros written)."""
        return self.state

    def _materialize(self):
        """Write |0...0> for real if it is still pending (see :meth:`StateVector.reset`)."""
        if getattr(self, "_fresh", False):
            self._fresh = False
            self._init_basis()

    def _run_big(self, s):
        self._materialize()
        torch = _torch()
        k = len(s.positions)
        m = np.ascontiguousarray(s.matrix, dtype=self.np_dtype)
        m_dev = torch.from_numpy(m.view(np.float32 if self.dtype == QB2_C64 else np.float64).reshape(-1)).to(self.device)
        tg = (ctypes.c_int * k)(*s.positions)
        dst = self._scratch_buf()
        _lib.check(
            self.lib.qb2_apply_matrix_big(
                self.state.data_ptr(), dst.data_ptr(), self.n, self.hi_base, m_dev.data_ptr(), tg, k, s.cmask, s.cval,
                self.dtype, self._stream(),
            ),
            "qb2_apply_matrix_big",
        )
        self.state, self._scratch = dst, self.state
        self.stats["big"] += 1


class StateVector(DeviceOps):
    """``num_qubits`` logical qubits in |0...0>.

    ``pos[q]`` is the physical position (bit of the array index) of logical qubit ``q``;
    the reference's canonical order is ``pos[q] = num_qubits - 1 - q``.
    """

    def __init__(self, num_qubits, dtype=np.complex64, device=None, T=None, r=None, L=4):
        torch = _torch()
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
        self.cfg = PlanConfig(self.n, self.dtype, T=T, r=r, L=L)
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
        """Back to |0...0>.  Lazy: the first pass that runs synthesises the zeros instead of reading
        them (one write and one read of the state saved); anything else that looks at the state
        first makes :meth:`_materialize` write them."""
        self._fresh = True

    def _init_basis(self):
        _lib.check(self.lib.qb2_init_basis(self._buf.data_ptr(), self.n, 0, self.dtype, self._stream()), "qb2_init_basis")

    def synchronize(self):
        _torch().cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ gates
    def compile(self, instructions):
        """Normalise + plan a list of unitary instructions against the CURRENT qubit map.
        Returns a compiled program that :meth:`execute` runs; the map is advanced."""
        # complex128 keeps non-unit diagonal entries exactly as given (see normalize.py)
        unit_tol = 1e-6 if self.dtype == QB2_C64 else 1e-13
        ops = normalize_circuit(instructions, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
        return self.compile_ops(ops)

    def compile_ops(self, ops):
        planner = Planner(self.cfg, self.pos)
        steps, need = planner.plan(ops)
        if need is not None:
            raise _lib.Qb2Error(f"op needs non-local positions {need} on a single-device state")
        return Program(self, steps)

    def execute(self, program):
        program.run()

    def apply_circuit(self, instructions):
        """Plan and run in one go, STREAMING: every pass is launched the moment the planner has built
        it (launches are asynchronous), so the host plans pass k+1 while the device sweeps pass k and
        planning disappears from the wall time except for the first pass.  Returns the step list."""
        unit_tol = 1e-6 if self.dtype == QB2_C64 else 1e-13
        ops = normalize_circuit(instructions, max_targets=self.cfg.max_mat, unit_tol=unit_tol)
        planner = Planner(self.cfg, self.pos)
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

        steps, need = planner.plan(ops, on_step=launch)
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
        cb = self.lib.qb2_sample_chunk_bits(self.n)
        sums = torch.empty(1 << (self.n - cb), dtype=torch.float64, device=self.device)
        _lib.check(
            self.lib.qb2_sample_prepare(self.state.data_ptr(), self.n, sums.data_ptr(), self.dtype, self._stream()),
            "qb2_sample_prepare",
        )
        u_dev = torch.from_numpy(uniforms).to(self.device)
        idx = torch.empty(shots, dtype=torch.int64, device=self.device)
        _lib.check(
            self.lib.qb2_sample_draw(
                self.state.data_ptr(), self.n, sums.data_ptr(), u_dev.data_ptr(), shots, idx.data_ptr(), self.dtype,
                self._stream(),
            ),
            "qb2_sample_draw",
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
    """A planned list of steps with its pass blobs resident on the device."""

    def __init__(self, sv, steps):
        self.sv = sv
        self.steps = steps
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
        self.n_passes = sum(1 for s in steps if s.kind == "pass")
        self.n_kernel_ops = sum(s.n_ops for s in steps if s.kind == "pass")
        self.n_gate_ops = sum(s.n_gate_ops for s in steps if s.kind == "pass") + sum(1 for s in steps if s.kind == "big")
        self.n_phases = sum(s.n_phases for s in steps if s.kind == "pass")

    def run(self):
        sv = self.sv
        for s, o, h in zip(self.steps, self.offsets, self._headers):
            if s.kind == "pass":
                sv._run_pass(s, h, self.dev, o)
            else:
                sv._run_big(s)
        sv.stats["passes"] += self.n_passes
        sv.stats["kernel_ops"] += self.n_kernel_ops
        sv.stats["gate_ops"] += self.n_gate_ops
        sv.stats["phases"] += self.n_phases


