"""Shared test helpers.  The CPU route (planner -> packed passes -> numpy emulator) checks
the host logic without a GPU; the GPU tests run the same blobs through libqb2.so."""

import numpy as np

from oracle import pass_emulator
from qrisp_b200.normalize import normalize_circuit
from qrisp_b200.planner import PlanConfig, Planner

MIN_QUBITS = 9


def plan_circuit(instrs, n, dtype=0, T=None, r=None, L=4, window=4096):
    """(steps, pos, N): plan the unitary instructions of an ``n``-qubit circuit."""
    N = max(n, MIN_QUBITS)
    cfg = PlanConfig(N, dtype=dtype, T=T, r=r, L=L)
    pos = [n - 1 - q for q in range(n)] + list(range(n, N))
    ops = normalize_circuit(instrs, max_targets=cfg.max_mat, unit_tol=1e-6 if dtype == 0 else 1e-13)
    steps, need = Planner(cfg, pos, window=window).plan(ops)
    assert need is None
    return steps, pos, N


def canonical(state, pos, n):
    """Physical-order state -> the reference's order (qubit 0 = MSB), first 2**n entries."""
    N = len(pos)
    j = np.arange(1 << N)
    i = np.zeros_like(j)
    for q in range(N):
        i |= ((j >> pos[q]) & 1) << (n - 1 - q if q < n else q)
    out = np.zeros_like(state)
    out[i] = state
    return out[: 1 << n]


def big_step_numpy(state, step):
    """Reference semantics of a BigMatStep on a physical-order numpy state."""
    N = int(np.log2(state.size))
    k = len(step.positions)
    idx = np.arange(1 << N)
    on = ((idx ^ step.cval) & step.cmask) == 0
    tmask = sum(1 << p for p in step.positions)
    row = np.zeros_like(idx)
    for t, p in enumerate(step.positions):
        row |= ((idx >> p) & 1) << t
    base = idx & ~tmask
    out = state.copy()
    acc = np.zeros(state.shape, dtype=np.complex128)
    for c in range(1 << k):
        dep = sum(((c >> t) & 1) << p for t, p in enumerate(step.positions))
        acc += step.matrix[row, c] * state[base | dep]
    out[on] = acc[on].astype(state.dtype)
    return out


def emulate_statevector(qc, dtype=0, report=None, sparse_start=False, max_support=None, **plan_kw):
    """Plan ``qc`` and run the packed passes through the numpy emulator.  ``sparse_start``: as the engine
    does it -- the host prefix (qb2_sparse_prefix) first, the first pass from the sparse input list."""
    from qrisp_b200 import prefix

    instrs = [i for i in qc.data if i.matrix is not None]
    n = qc.num_qubits
    cdt = np.complex64 if dtype == 0 else np.complex128
    pending = None
    if sparse_start:
        consumed, pending = prefix.sparse_prefix(instrs, prefix.zero_state(), max_support)
        instrs = instrs[consumed:]
    steps, pos, N = plan_circuit(instrs, n, dtype=dtype, **plan_kw)
    pos0 = [n - 1 - q for q in range(n)] + list(range(n, N))  # the map the first step sees
    st = np.zeros(1 << N, dtype=cdt)
    if pending is None:
        st[0] = 1
    elif not steps or steps[0].kind != "pass":
        st[prefix.to_physical(pending[0], pos0, n).astype(np.int64)] = pending[1].astype(cdt)
        pending = None
    for s in steps:
        if s.kind == "pass":
            sparse = None
            if pending is not None:
                phys = prefix.to_physical(pending[0], pos0, n)
                data, k = prefix.sparse_input_list(s, phys, pending[1], N, 0, cdt)
                pad = (4 * k + 15) // 16 * 16
                sparse = (np.frombuffer(data, np.uint32, k, 0), np.frombuffer(data, np.uint32, k, pad), np.frombuffer(data, cdt, k, 2 * pad))
                st = np.full(1 << N, np.nan, dtype=cdt)  # never read
                pending = None
            st = pass_emulator.run(s.blob, st, report=report, sparse=sparse)
        else:
            st = big_step_numpy(st, s)
    return canonical(st, pos, n), steps


def marginal(state, qubits, n):
    psi = state.reshape(n * [2])
    m = len(qubits)
    psi = np.moveaxis(psi, list(qubits), list(range(m))).reshape(1 << m, -1)
    return (np.abs(psi.astype(np.complex128)) ** 2).sum(axis=1)
