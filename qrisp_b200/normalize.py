"""Gate normalisation: circuit instructions -> the three op shapes the B200 pass kernel runs.

The reference applies every instruction as a dense ``2**k x 2**k`` matrix
(reference: quantum_state.py:41-78 -> tensor_factor.py:67-89).  A streaming GPU kernel can do
much better when it knows the *structure* of the matrix, so every instruction is classified
once, on the host, from its unitary:

* :class:`DiagTerm` -- the unitary is diagonal.  It becomes a phase table; any number of
  diagonal gates on any qubits (tile, non-tile, other ranks) are merged into one table
  product.
* :class:`CMat` -- a dense matrix on a few *target* qubits, applied when a set of *control*
  qubits has a given value.  A qubit is a control when the unitary is block diagonal in it
  (it commutes with Z on that qubit) -- the same criterion the reference's permeability
  analysis uses (reference: permeability/type_checker.py ``is_permeable``).  Controls cost
  nothing: they may sit anywhere in the index.
* :class:`Relabel` -- an exact SWAP: no data moves, the qubit -> position map changes
  (reference: TensorFactor.swap, tensor_factor.py:56-64, which *does* transpose the data).

Index convention of the normalised ops: LITTLE endian in their own qubit list (bit ``i`` of a
matrix / table index is the value of ``qubits[i]``).  The instruction convention is the
reference's big endian (circuit.py).
"""

from __future__ import annotations

import numpy as np

TOL = 1e-10
MAX_DIAG_QUBITS = 6  # widest diagonal gate that becomes one phase table

_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


class DiagTerm:
    """``amp[x] *= exp(2 pi i * turns[bits of x at qubits])``."""

    __slots__ = ("qubits", "turns")
    kind = "diag"

    def __init__(self, qubits, turns):
        self.qubits = tuple(int(q) for q in qubits)
        self.turns = np.asarray(turns, dtype=np.float64)
        assert self.turns.shape == (1 << len(self.qubits),)

    @property
    def nd(self):  # qubits acted on non-diagonally
        return ()

    @property
    def d(self):  # qubits acted on diagonally
        return self.qubits

    def matrix(self):
        return np.diag(np.exp(2j * np.pi * self.turns))

    def __repr__(self):
        return f"DiagTerm({self.qubits})"


class CMat:
    """Dense ``matrix`` on ``targets`` (little endian) when every ``controls[q] == value``."""

    __slots__ = ("targets", "matrix", "controls")
    kind = "mat"

    def __init__(self, targets, matrix, controls=None):
        self.targets = tuple(int(q) for q in targets)
        self.matrix = np.ascontiguousarray(matrix, dtype=np.complex128)
        self.controls = dict(controls or {})
        d = 1 << len(self.targets)
        assert self.matrix.shape == (d, d)

    @property
    def nd(self):
        return self.targets

    @property
    def d(self):
        return tuple(self.controls.keys())

    def __repr__(self):
        return f"CMat(targets={self.targets}, controls={self.controls})"


class Relabel:
    """Exchange the physical positions of two logical qubits (an exact SWAP gate)."""

    __slots__ = ("a", "b")
    kind = "relabel"

    def __init__(self, a, b):
        self.a, self.b = int(a), int(b)

    @property
    def nd(self):
        return (self.a, self.b)

    @property
    def d(self):
        return ()

    def __repr__(self):
        return f"Relabel({self.a}, {self.b})"


def _bit(x, i):
    return (x >> i) & 1


def diagonal_qubits(u, k):
    """Bits ``i`` of the (little endian) index in which ``u`` is block diagonal."""
    dim = 1 << k
    nz_r, nz_c = np.nonzero(np.abs(u) > TOL)
    diff = np.bitwise_xor(nz_r, nz_c)
    mixed = int(np.bitwise_or.reduce(diff)) if diff.size else 0
    return [i for i in range(k) if not _bit(mixed, i)], [i for i in range(k) if _bit(mixed, i)]


_TEMPLATES = {}  # (matrix bytes, k, max_targets, max_blocks, unit_tol) -> ops on qubits 0..k-1


def normalize_instruction(matrix, qubits, max_targets=4, max_blocks=4, unit_tol=1e-6):
    """Classify one unitary instruction.  ``matrix`` is big endian in ``qubits``.

    Returns a list of normalised ops that, applied in order, equal the instruction.
    A dense block on more than ``max_targets`` qubits is returned as a :class:`CMat`
    all the same; the planner routes it to the big-matrix path.

    Circuits repeat a handful of matrices thousands of times, so the classification is
    memoised on the matrix bytes and re-targeted to the instruction's qubits.
    """
    k = len(qubits)
    if k <= 3:
        u = np.ascontiguousarray(matrix, dtype=np.complex128)
        key = (u.tobytes(), k, max_targets, max_blocks, unit_tol)
        tmpl = _TEMPLATES.get(key)
        if tmpl is None:
            if len(_TEMPLATES) > 4096:
                _TEMPLATES.clear()
            tmpl = _normalize_instruction(u, tuple(range(k)), max_targets, max_blocks, unit_tol)
            _TEMPLATES[key] = tmpl
        out = []
        for op in tmpl:
            if op.kind == "relabel":
                out.append(Relabel(qubits[op.a], qubits[op.b]))
            elif op.kind == "diag":
                out.append(DiagTerm([qubits[q] for q in op.qubits], op.turns))
            else:
                out.append(CMat([qubits[q] for q in op.targets], op.matrix, {qubits[q]: v for q, v in op.controls.items()}))
        return out
    return _normalize_instruction(matrix, qubits, max_targets, max_blocks, unit_tol)


def _normalize_instruction(matrix, qubits, max_targets, max_blocks, unit_tol):
    k = len(qubits)
    u = np.asarray(matrix, dtype=np.complex128)
    q_le = tuple(qubits[::-1])  # index bit i <-> q_le[i]
    if k == 2 and np.max(np.abs(u - _SWAP)) < TOL:
        return [Relabel(qubits[0], qubits[1])]
    dbits, tbits = diagonal_qubits(u, k)
    if not tbits and np.max(np.abs(np.abs(np.diagonal(u)) - 1.0)) > unit_tol:
        # diagonal entries that are not of unit modulus to the engine's precision (matrices
        # that were rounded to complex64): a phase table would silently renormalise them, a
        # 2x2 diagonal block keeps the values as given
        dbits, tbits = list(range(1, k)), [0]
    if not tbits and k > MAX_DIAG_QUBITS:
        # a wide diagonal gate (e.g. a many-controlled phase): one qubit plays the target of
        # 2x2 diagonal blocks, the others are controls -- usually a single non-identity block
        dbits, tbits = list(range(1, k)), [0]
    if not tbits:
        diag = np.diagonal(u)
        turns = np.angle(diag) / (2 * np.pi)
        if np.max(np.abs(diag - 1.0)) < 1e-15:
            return []
        # drop qubits the phases do not depend on
        t = turns.reshape(k * [2])  # axis a <-> bit k-1-a
        keep = []
        for i in range(k):
            ax = k - 1 - i
            lo = np.take(t, 0, axis=ax)
            hi = np.take(t, 1, axis=ax)
            d = np.exp(2j * np.pi * lo) - np.exp(2j * np.pi * hi)
            if np.max(np.abs(d)) > 1e-15:
                keep.append(i)
        if len(keep) < k:
            if not keep:
                keep = [0]
            idx = np.zeros(1 << len(keep), dtype=np.int64)
            for pos, i in enumerate(keep):
                idx |= ((np.arange(1 << len(keep)) >> pos) & 1) << i
            return [DiagTerm([q_le[i] for i in keep], turns[idx])]
        return [DiagTerm(q_le, turns)]
    nt, ndg = len(tbits), len(dbits)
    # index helpers
    tvals = np.zeros(1 << nt, dtype=np.int64)
    for pos, i in enumerate(tbits):
        tvals |= ((np.arange(1 << nt) >> pos) & 1) << i
    blocks = []
    for x in range(1 << ndg):
        off = 0
        for pos, i in enumerate(dbits):
            off |= _bit(x, pos) << i
        rows = tvals + off
        blk = u[np.ix_(rows, rows)]
        if np.max(np.abs(blk - np.eye(1 << nt))) < 1e-14:
            continue
        blocks.append((x, blk))
    if not blocks:
        return []
    if len(blocks) > max_blocks and k <= max_targets:
        return [CMat(q_le, u)]
    out = []
    for x, blk in blocks:
        ctrl = {q_le[i]: _bit(x, pos) for pos, i in enumerate(dbits)}
        out.append(CMat([q_le[i] for i in tbits], blk, ctrl))
    return out


def normalize_circuit(instructions, max_targets=4, unit_tol=1e-6):
    """Normalise a list of unitary ``Instruction`` objects (circuit.py) and fuse what is
    free to fuse on the host:

    * runs of uncontrolled one-qubit gates on the same qubit are multiplied together,
    * a pending one-qubit gate is folded into the next uncontrolled dense gate that has the
      qubit as a target, and into the previous one when nothing touched the qubit since.

    This is the part of the reference's ``group_qc`` (circuit_preprocessing.py:119-157) that
    pays off here: bigger fused unitaries would turn a bandwidth-bound sweep into a
    compute-bound one (a dense k-qubit block costs 4*2**k FMA per amplitude).
    """
    out = []
    pending = {}  # qubit -> 2x2 matrix not yet emitted
    open_dense = {}  # qubit -> index in `out` of an uncontrolled dense CMat nothing has touched since

    def flush(q):
        m = pending.pop(q, None)
        if m is not None:
            if np.max(np.abs(m - np.eye(2))) > 1e-15:
                for op in normalize_instruction(m, (q,), max_targets, unit_tol=unit_tol):
                    out.append(op)
            open_dense.pop(q, None)

    for instr in instructions:
        ops = normalize_instruction(instr.matrix, instr.qubits, max_targets, unit_tol=unit_tol)
        for op in ops:
            if op.kind == "diag" and len(op.qubits) == 1 and (op.qubits[0] in pending or op.qubits[0] in open_dense):
                # a one-qubit phase next to a dense gate on the same qubit: multiply it in
                op = CMat(op.qubits, op.matrix())
            if op.kind == "mat" and not op.controls and len(op.targets) == 1:
                q = op.targets[0]
                idx = open_dense.get(q)
                if idx is not None and q not in pending:
                    host = out[idx]
                    i = host.targets.index(q)
                    host.matrix = _embed(op.matrix, i, len(host.targets)) @ host.matrix
                else:
                    pending[q] = op.matrix @ pending[q] if q in pending else op.matrix
                continue
            if op.kind == "mat" and not op.controls and len(op.targets) <= max_targets:
                m = op.matrix
                for i, q in enumerate(op.targets):
                    p = pending.pop(q, None)
                    if p is not None:
                        m = m @ _embed(p, i, len(op.targets))
                op.matrix = m
                for q in op.d:
                    flush(q)
                    open_dense.pop(q, None)
                out.append(op)
                for q in op.targets:
                    open_dense[q] = len(out) - 1
                continue
            for q in tuple(op.nd) + tuple(op.d):
                flush(q)
                open_dense.pop(q, None)
            out.append(op)
    for q in list(pending):
        flush(q)
    return out


def _embed(m2, i, k):
    """2x2 ``m2`` on bit ``i`` of a ``k``-bit little-endian index."""
    res = np.eye(1, dtype=np.complex128)
    for b in range(k - 1, -1, -1):  # most significant first for kron
        res = np.kron(res, m2 if b == i else np.eye(2))
    return res
