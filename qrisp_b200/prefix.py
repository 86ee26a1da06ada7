"""Sparse start of a simulation: host-side evolution of a state with few non-zero amplitudes.

The reference keeps a fresh qubit as a sparse ``BiArray`` and contracts gates sparsely while the
state is sparse (reference: tensor_factor.py:32-50, :86-89; bi_arrays.py:1197-1213) -- which is why
a GHZ preparation costs it next to nothing.  The B200 engine does the same in front of its dense
sweeps: ``qb2_sparse_prefix`` (C++, host only, include/qrisp_b200.h) walks the unitary instructions
for as long as the list of non-zero amplitudes stays within ``max_support`` entries; the first dense
pass then builds its tiles from that list instead of reading a state (QB2_FEATURE_ZERO_INPUT), and
tiles without an entry are written as zeros without running any op.

A state is a pair ``(idx, amp)``: ``idx`` uint64, bit ``q`` = value of logical qubit ``q``; ``amp``
complex128; sorted by index.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib

MAX_SUPPORT = 1024  # QB2_MAX_SPARSE
CHUNK = 96  # instructions packed per call: the walk usually ends within the first chunks
DROP_TOL = 1e-14


def default_support():
    try:
        return max(0, min(MAX_SUPPORT, int(os.environ.get("QB2_PREFIX_SUPPORT", MAX_SUPPORT))))
    except ValueError:
        return MAX_SUPPORT


def zero_state():
    return np.zeros(1, dtype=np.uint64), np.ones(1, dtype=np.complex128)


def is_zero_state(state):
    idx, amp = state
    return len(idx) == 1 and int(idx[0]) == 0 and amp[0] == 1.0


def sparse_prefix(instructions, state, max_support=None):
    """Apply a prefix of ``instructions`` (unitary :class:`~qrisp_b200.circuit.Instruction` objects) to
    the sparse ``state``.  Returns ``(consumed, state)``."""
    max_support = default_support() if max_support is None else int(max_support)
    idx0, amp0 = state
    if max_support < 1 or len(idx0) > max_support or not instructions:
        return 0, state
    lib = _lib.load()
    idx = np.zeros(max_support, dtype=np.uint64)
    amp = np.zeros(max_support, dtype=np.complex128)
    idx[: len(idx0)] = idx0
    amp[: len(idx0)] = amp0
    count = ctypes.c_uint32(len(idx0))
    consumed = ctypes.c_int32(0)
    total = 0
    while total < len(instructions):
        chunk = instructions[total : total + CHUNK]
        usable = 0
        for ins in chunk:  # stop in front of anything the walk cannot take
            if ins.matrix is None or len(ins.qubits) > 12:
                break
            usable += 1
        if usable == 0:
            break
        chunk = chunk[:usable]
        qoff = np.zeros(len(chunk) + 1, dtype=np.int32)
        moff = np.zeros(len(chunk), dtype=np.int64)
        qidx, mats, mo = [], [], 0
        for i, ins in enumerate(chunk):
            qidx.extend(ins.qubits)
            qoff[i + 1] = len(qidx)
            moff[i] = mo
            mats.append(ins.matrix.ravel())
            mo += ins.matrix.size
        qidx = np.asarray(qidx, dtype=np.int32)
        mat = np.ascontiguousarray(np.concatenate(mats), dtype=np.complex128)
        _lib.check(
            lib.qb2_sparse_prefix(
                len(chunk), qoff.ctypes.data, qidx.ctypes.data, moff.ctypes.data, mat.ctypes.data, max_support, DROP_TOL,
                idx.ctypes.data, amp.ctypes.data, ctypes.byref(count), ctypes.byref(consumed),
            ),
            "qb2_sparse_prefix",
        )
        total += consumed.value
        if consumed.value < len(chunk) or usable < CHUNK:
            break
    k = count.value
    return total, (idx[:k].copy(), amp[:k].copy())


def to_physical(idx, pos, n_logical):
    """Logical indices (bit q = qubit q) -> physical indices (bit pos[q])."""
    out = np.zeros(idx.shape, dtype=np.uint64)
    one = np.uint64(1)
    for q in range(n_logical):
        out |= ((idx >> np.uint64(q)) & one) << np.uint64(pos[q])
    return out


def sparse_input_list(step, phys, amp, n_local, hi_base, np_dtype):
    """The list ``qb2_run_pass`` takes for a pass that runs on a never-written state: the entries of this
    rank (positions >= n_local spell ``hi_base``), as uint32 tile numbers (ascending), uint32
    ``(thread << r) | register`` under the pass's first phase, and amplitudes.  Returns
    ``(bytes, count)``: the three arrays, each padded to 16 bytes."""
    one = np.uint64(1)
    if len(phys):
        mine = (phys >> np.uint64(n_local)) == np.uint64(hi_base >> n_local)
        phys, amp = phys[mine], amp[mine]
    local = phys & np.uint64((1 << n_local) - 1)
    tile = np.zeros(local.shape, dtype=np.uint64)
    for src, dst, ln in step.segs:
        tile |= ((local >> np.uint64(dst)) & np.uint64((1 << ln) - 1)) << np.uint64(src)
    thread = np.zeros(local.shape, dtype=np.uint64)
    for b, x in enumerate(step.TP0):
        thread |= ((local >> np.uint64(x)) & one) << np.uint64(b)
    reg = np.zeros(local.shape, dtype=np.uint64)
    for i, x in enumerate(step.R0):
        reg |= ((local >> np.uint64(x)) & one) << np.uint64(i)
    order = np.argsort(tile, kind="stable")
    k = len(order)
    pad = (4 * k + 15) // 16 * 16
    buf = bytearray(2 * pad + k * np.dtype(np_dtype).itemsize)
    buf[0 : 4 * k] = tile[order].astype(np.uint32).tobytes()
    buf[pad : pad + 4 * k] = ((thread[order] << np.uint64(len(step.R0))) | reg[order]).astype(np.uint32).tobytes()
    buf[2 * pad :] = amp[order].astype(np_dtype).tobytes()
    return bytes(buf), k
