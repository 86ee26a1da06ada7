"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A plain-numpy restatement of the reference's statevector path
(``qrisp.simulator.run`` / ``statevector_sim``), used as the checker for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; nothing under ``qrisp_b200/`` does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
fixtures under ``tests/golden/`` that were produced by running the reference itself
(``tests/golden/make_golden.py`` imports ``/root/reference/src/qrisp`` with import stubs
for the absent jax / qiskit / matplotlib packages and records ``statevector_sim`` and
``run`` outputs).

The reference keeps the state as lazily entangled ``TensorFactor`` objects over
dense/sparse ``BiArray``s; mathematically that is one statevector, so the oracle keeps
one dense numpy array and restates the numerics that reach the result:

* gate application           -- tensor_factor.py:67-89   (``apply_matrix``)
* the fused-matrix threshold -- tensor_factor.py:72-73   (entries < float_thresh zeroed)
* marginal probabilities     -- bi_arrays.py:1030-1055, bi_array_helper.py:315-387
* outcome labelling          -- simulator.py:151-203, quantum_state.py:174-226
* allocation / reset / mid-circuit measurement rewriting
                             -- circuit_preprocessing.py:466-491, :729-836
"""

from __future__ import annotations

import numpy as np

# reference: numerics_config.py:24-38
FLOAT_THRESH = np.float32(1e-5)
CUTOFF_RATIO = np.float32(2e-4)


# ------------------------------------------------------------------------------------
# gate application
# ------------------------------------------------------------------------------------


def apply_matrix(state, matrix, qubits, n=None, threshold=True):
    """Apply ``matrix`` (big-endian in ``qubits``) to ``state`` (qubit 0 = MSB).

    reference: tensor_factor.py:67-89 -- move ``qubits[i]`` to axis ``i``, reshape to
    ``(2**k, 2**(n-k))``, contract ``matrix @ state``.  The reference leaves the axes
    permuted and tracks the order in ``TensorFactor.qubits``; ``unravel``
    (tensor_factor.py:163-169) sorts them again, which is what this function returns.
    """
    if n is None:
        n = int(np.log2(state.size))
    k = len(qubits)
    matrix = np.asarray(matrix)
    if threshold and matrix.size >= 2**6:
        # reference: tensor_factor.py:72-73
        matrix = matrix * (np.abs(matrix) > FLOAT_THRESH)
    psi = state.reshape(n * [2])
    psi = np.moveaxis(psi, list(qubits), list(range(k)))
    shape = psi.shape
    psi = matrix.astype(state.dtype) @ psi.reshape(1 << k, -1)
    psi = np.moveaxis(psi.reshape(shape), list(range(k)), list(qubits))
    return np.ascontiguousarray(psi).reshape(-1)


def zero_state(n, dtype=np.complex64):
    psi = np.zeros(1 << n, dtype=dtype)
    psi[0] = 1
    return psi


def statevector(circuit, dtype=np.complex64, threshold=False):
    """reference: simulator.py:220-296 ``statevector_sim`` (no measurements allowed;
    qb_alloc / qb_dealloc / barrier are dropped, circuit_preprocessing.py:466-491 with
    ``insert_reset=False``)."""
    n = circuit.num_qubits
    psi = zero_state(n, dtype)
    for instr in circuit.data:
        if instr.name in ("barrier", "qb_alloc", "qb_dealloc"):
            continue
        if instr.name == "measure":
            raise Exception("Tried to determine the statevector of a circuit containing a measurement")
        if instr.matrix is None:
            raise Exception(f"statevector: non-unitary instruction {instr.name!r}")
        psi = apply_matrix(psi, instr.matrix, instr.qubits, n, threshold=threshold)
    return psi


# ------------------------------------------------------------------------------------
# measurement
# ------------------------------------------------------------------------------------


def marginal_probabilities(state, mes_qubits, n=None):
    """``p[o]`` for every outcome ``o`` of ``mes_qubits`` with ``mes_qubits[0]`` as the
    most significant outcome bit (reference: tensor_factor.py:171-182 swaps
    ``mes_qubits[i]`` to axis ``i``; bi_array_helper.py:315-327 takes row norms of the
    ``(2**m, 2**(n-m))`` reshape)."""
    if n is None:
        n = int(np.log2(state.size))
    m = len(mes_qubits)
    psi = state.reshape(n * [2])
    psi = np.moveaxis(psi, list(mes_qubits), list(range(m))).reshape(1 << m, -1)
    return np.einsum("ij,ij->i", psi.real.astype(np.float64), psi.real.astype(np.float64)) + np.einsum(
        "ij,ij->i", psi.imag.astype(np.float64), psi.imag.astype(np.float64)
    )


def multi_measure(state, mes_qubits, n=None):
    """Outcome indices and probabilities that survive the reference's pruning.

    reference: bi_arrays.py:1030-1055 -- more than 10 measured qubits use
    ``dense_measurement_brute`` (keep ``p > max(p) * cutoff_ratio``,
    bi_array_helper.py:315-331); otherwise ``dense_measurement_smart`` recursively halves
    the array and drops branches whose norm is below ``float_thresh``
    (bi_array_helper.py:349-387).  tensor_factor.py:184 and quantum_state.py:224 then
    normalise.
    """
    p = marginal_probabilities(state, mes_qubits, n)
    m = len(mes_qubits)
    if m > 10:
        keep = np.nonzero(p > p.max() * CUTOFF_RATIO)[0]
    else:
        # a branch is pruned when the norm of ANY enclosing half falls below float_thresh
        alive = np.ones(1, dtype=bool)
        for level in range(m + 1):
            width = 1 << (m - level)
            branch_p = p.reshape(1 << level, width).sum(axis=1)
            alive = alive & ~(branch_p < FLOAT_THRESH)
            if level < m:
                alive = np.repeat(alive, 2)
        keep = np.nonzero(alive)[0]
    probs = p[keep]
    probs = probs / probs.sum()
    return keep.astype(np.int64), probs


def finalize_probabilities(cl_prob):
    """reference: simulator.py:167-169 -- round to 5 significant digits relative to the
    largest probability, renormalise."""
    cl_prob = np.asarray(cl_prob, dtype=np.float64)
    cl_prob = np.round(cl_prob, int(-np.log10(np.max(cl_prob))) + 5)
    return cl_prob / np.sum(cl_prob)


# ------------------------------------------------------------------------------------
# run(): circuit rewriting + measurement labelling
# ------------------------------------------------------------------------------------


def _is_permeable(instr, qubit):
    """reference: permeability/type_checker.py ``is_permeable`` -- an operation is
    permeable on a qubit when its unitary commutes with Z on that qubit, i.e. it is
    block diagonal in that qubit's computational basis."""
    if instr.matrix is None:
        return False
    k = len(instr.qubits)
    pos = instr.qubits.index(qubit)
    t = instr.matrix.reshape(2 * k * [2])
    t = np.moveaxis(t, [pos, k + pos], [0, 1])
    return bool(np.all(np.abs(t[0, 1]) < 1e-10) and np.all(np.abs(t[1, 0]) < 1e-10))


def rewrite_measurements(circuit):
    """Return ``(unitary_instructions, num_qubits, measurements)`` where mid-circuit
    measurements and resets have been replaced by CX onto fresh qubits.

    reference: circuit_preprocessing.py:729-836 ``insert_multiverse_measurements``.  The
    reference also drops ``disentangle`` hints into the stream; they never change the state
    (tensor_factor.py:198-247 only splits off a qubit that already is in a product state),
    so the oracle omits them.  A reset of qubit q becomes CX(q, fresh); CX(fresh, q):
    q ends in |0> in every branch and the branch information moves to ``fresh``.
    """
    CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    n = circuit.num_qubits
    data = [i for i in circuit.data if i.name not in ("barrier", "qb_alloc")]
    out = []
    measurements = []  # (qubit, clbit)
    cb_to_qb = {}
    j = 0
    from qrisp_b200.circuit import Instruction  # plain container, no compute

    while j < len(data):
        instr = data[j]
        j += 1
        if instr.name == "measure":
            q, c = instr.qubits[0], instr.clbits[0]
            followed_by_reset = False
            terminal = True
            for l in range(j, len(data)):
                nxt = data[l]
                if q in nxt.qubits:
                    if nxt.name == "reset":
                        followed_by_reset = True
                        data.pop(l)
                        terminal = False
                        break
                    if nxt.name in ("qb_dealloc", "disentangle"):
                        continue
                    if not _is_permeable(nxt, q):
                        terminal = False
                        break
                if nxt.name == "measure" and nxt.qubits[0] == q:
                    terminal = False
                    break
                if c in nxt.clbits:
                    terminal = False
                    break
            if terminal:
                measurements.append((q, c))
                continue
            fresh = n
            n += 1
            out.append(Instruction("cx", (q, fresh), (), CX))
            if followed_by_reset:
                out.append(Instruction("cx", (fresh, q), (), CX))
            cb_to_qb[c] = fresh
        elif instr.name == "reset":
            q = instr.qubits[0]
            needed = False
            for l in range(j, len(data)):
                nxt = data[l]
                if q in nxt.qubits and nxt.name not in ("qb_dealloc", "disentangle") and not _is_permeable(nxt, q):
                    needed = True
                    break
            if not needed:
                continue
            fresh = n
            n += 1
            out.append(Instruction("cx", (q, fresh), (), CX))
            out.append(Instruction("cx", (fresh, q), (), CX))
        elif instr.name in ("qb_dealloc", "disentangle"):
            continue
        elif getattr(instr, "ctrl_state", None) is not None:
            # reference: circuit_preprocessing.py:790-816 (ClControlledOperation): controls on the qubits
            # that carry the outcomes; an unwritten clbit with ctrl_state "0" gets a fresh |0> qubit, with
            # "1" the operation is dropped
            ctrl_qubits = []
            for cb, want in zip(instr.clbits, instr.ctrl_state):
                if cb not in cb_to_qb:
                    if want == "1":
                        break
                    ctrl_qubits.append(n)
                    n += 1
                else:
                    ctrl_qubits.append(cb_to_qb[cb])
            else:
                k, nb = len(ctrl_qubits), instr.matrix.shape[0]
                m = np.eye(nb << k, dtype=np.complex128)
                cs = int(instr.ctrl_state, 2) if k else 0
                m[cs * nb : (cs + 1) * nb, cs * nb : (cs + 1) * nb] = instr.matrix  # unitary_management.py:68-88
                out.append(Instruction(instr.name, tuple(ctrl_qubits) + tuple(instr.qubits), (), m))
        else:
            out.append(instr)
    for c, q in cb_to_qb.items():
        measurements.append((q, c))
    measurements = list(dict.fromkeys(measurements))
    return out, n, measurements


def run_distribution(circuit, dtype=np.complex64):
    """``(outcome_strings, probabilities)`` of ``qrisp.simulator.run(qc, shots=None)``.

    reference: simulator.py:44-203.  Measurements are ordered by descending clbit index
    (simulator.py:151), the resulting outcome integer is printed MSB first, so clbit 0 is
    the right-most character (qiskit order).
    """
    if len(circuit.data) == 0:
        return [""], np.array([1.0])
    if not any(i.name == "measure" for i in circuit.data):
        return [""], np.array([1.0])
    instrs, n, measurements = rewrite_measurements(circuit)
    psi = zero_state(n, dtype)
    for instr in instrs:
        psi = apply_matrix(psi, instr.matrix, instr.qubits, n, threshold=False)
    measurements.sort(key=lambda x: -x[1])
    mes_qubits = [q for q, _ in measurements]
    outcomes, probs = multi_measure(psi, mes_qubits, n)
    probs = finalize_probabilities(probs)
    width = len(measurements)
    acc = {}
    for o, p in zip(outcomes, probs):
        if p == 0:
            continue
        s = bin(int(o))[2:].zfill(width)
        acc[s] = acc.get(s, 0.0) + float(p)
    keys = list(acc.keys())
    return keys, np.array([acc[k] for k in keys])


def run(circuit, shots=None, seed=None, dtype=np.complex64):
    """Counts dict of ``qrisp.simulator.run`` (reference: simulator.py:171-203)."""
    if len(circuit.data) == 0:
        return {"": shots}
    if shots == 0:
        return {}
    keys, probs = run_distribution(circuit, dtype)
    if keys == [""]:
        return {"": 1.0}
    if shots is None:
        return dict(zip(keys, (float(p) for p in probs)))
    rng = np.random.default_rng(seed)
    samples = rng.choice(len(probs), shots, p=probs)
    hist = np.bincount(samples, minlength=len(probs))
    return {keys[i]: int(hist[i]) for i in range(len(keys)) if hist[i]}


# ------------------------------------------------------------------------------------
# fused unitaries
# ------------------------------------------------------------------------------------


def circuit_unitary(instrs, qubits):
    """Unitary of a list of instructions restricted to ``qubits`` (big-endian in that
    list).  reference: unitary_management.py:286-354 ``calc_circuit_unitary`` -- the
    product of the embedded gate unitaries, later gates multiplied from the left."""
    k = len(qubits)
    pos = {q: i for i, q in enumerate(qubits)}
    u = np.eye(1 << k, dtype=np.complex128)
    for instr in instrs:
        loc = [pos[q] for q in instr.qubits]
        # apply to every column of u
        cols = u.reshape(k * [2] + [1 << k])
        kk = len(loc)
        cols = np.moveaxis(cols, loc, list(range(kk)))
        shape = cols.shape
        cols = instr.matrix @ cols.reshape(1 << kk, -1)
        u = np.moveaxis(cols.reshape(shape), list(range(kk)), loc).reshape(1 << k, 1 << k)
    return u


# ------------------------------------------------------------------------------------
# grouped application (what makes the reference fast on a CPU)
# ------------------------------------------------------------------------------------


def group_instructions(instrs, kmax=6):
    """Greedy restatement of the EFFECT of the reference's ``group_qc``
    (circuit_preprocessing.py:119-157): consecutive instructions are merged while their joint
    support stays within ``kmax`` qubits; each group is applied as ONE dense unitary
    (simulator.py:255 -> GroupedInstruction.get_instruction -> to_op().get_unitary()).  The
    reference searches for the grouping with the best flop gain; the greedy rule below reaches
    comparable group sizes on the benchmark circuits and is used for the CPU baseline only.
    Returns a list of ``(qubits, unitary)``."""
    groups = []
    cur, qubits = [], []
    for instr in instrs:
        new = [q for q in instr.qubits if q not in qubits]
        if cur and len(qubits) + len(new) > kmax:
            groups.append((list(qubits), circuit_unitary(cur, qubits)))
            cur, qubits = [], []
            new = list(instr.qubits)
        qubits.extend(new)
        cur.append(instr)
    if cur:
        groups.append((list(qubits), circuit_unitary(cur, qubits)))
    return groups


def statevector_grouped(circuit, dtype=np.complex64, kmax=6):
    """``statevector`` through :func:`group_instructions` (same result, fewer sweeps)."""
    n = circuit.num_qubits
    instrs = [i for i in circuit.data if i.matrix is not None]
    psi = zero_state(n, dtype)
    for qubits, u in group_instructions(instrs, kmax):
        psi = apply_matrix(psi, u, qubits, n, threshold=False)
    return psi


# ------------------------------------------------------------------------------------
# sparse application (the reference's SparseBiArray path)
# ------------------------------------------------------------------------------------


def _result_density(da, db, contraction):
    """reference: bi_arrays.py:1197-1204 -- estimated density of a contraction result
    (the reference calls the fraction of non-zeros ``sparsity``)."""
    return 1.0 - (1.0 - da * db) ** contraction


def apply_matrix_sparse(idx, val, matrix, qubits, n):
    """``matrix`` applied to the sparse state ``(idx, val)`` (sorted basis indices, amplitudes).

    Restates what the reference's sparse contraction computes (bi_arrays.py:415-520
    ``SparseBiArray.contract`` via bi_array_helper's coordinate matching): amplitudes are
    grouped by the index bits outside ``qubits``, every group is multiplied by ``matrix`` and
    exact zeros of the product are dropped, so a diagonal gate keeps the support and a
    Hadamard at most doubles it."""
    k = len(qubits)
    shifts = [n - 1 - q for q in qubits]  # qubit 0 is the most significant index bit
    tmask = 0
    for s in shifts:
        tmask |= 1 << s
    local = np.zeros(idx.shape, dtype=np.int64)
    for i, s in enumerate(shifts):
        local |= ((idx >> s) & 1) << (k - 1 - i)
    groups, inv = np.unique(idx & ~tmask, return_inverse=True)
    dense = np.zeros((1 << k, groups.size), dtype=val.dtype)
    dense[local, inv] = val
    out = np.asarray(matrix).astype(val.dtype) @ dense
    rows, cols = np.nonzero(out)
    new_idx = groups[cols]
    for i, s in enumerate(shifts):
        new_idx = new_idx | (((rows >> (k - 1 - i)) & 1) << s)
    new_val = out[rows, cols]
    order = np.argsort(new_idx, kind="stable")
    return new_idx[order], new_val[order]


def statevector_reference_path(circuit, dtype=np.complex64, kmax=6, sparsity_threshold=0.05):
    """The statevector the way the reference reaches it FAST: grouped unitaries
    (:func:`group_instructions`) applied to a state that starts sparse
    (tensor_factor.py:40-50) and is contracted sparsely while the estimated density of the
    result stays below ``contract_sparsity_threshold`` = 0.05 (tensor_factor.py:86-89,
    bi_arrays.py:1197-1213), densely afterwards.  This is the CPU arm of ``bench.py``: on
    GHZ-like inputs the sparse prefix is what keeps the reference's cost near O(2**n)."""
    n = circuit.num_qubits
    instrs = [i for i in circuit.data if i.matrix is not None]
    idx = np.zeros(1, dtype=np.int64)
    val = np.ones(1, dtype=dtype)
    psi = None
    size = float(1 << n)
    for qubits, u in group_instructions(instrs, kmax):
        if psi is None:
            dm = np.count_nonzero(np.abs(u) > FLOAT_THRESH) / u.size
            dense_next = _result_density(dm, idx.size / size, 1 << len(qubits)) > sparsity_threshold or size * u.size < 2**12
            if dense_next:
                psi = np.zeros(1 << n, dtype=dtype)
                psi[idx] = val
        if psi is None:
            idx, val = apply_matrix_sparse(idx, val, u * (np.abs(u) > FLOAT_THRESH), qubits, n)
        else:
            psi = apply_matrix(psi, u, qubits, n, threshold=True)
    if psi is None:
        psi = np.zeros(1 << n, dtype=dtype)
        psi[idx] = val
    return psi
