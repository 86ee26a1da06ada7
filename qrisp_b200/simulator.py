"""``run`` / ``statevector_sim``: the reference-facing entry points of the B200 path.

Same names, argument meaning and result contract as ``qrisp.simulator.run(qc, shots, token)``
and ``qrisp.simulator.statevector_sim(qc)`` (reference: simulator.py:44-203 and :220-296):
a circuit goes in, a counts / probability dict keyed by bitstrings (clbit 0 right-most)
comes out.  ``qc`` may be a :class:`qrisp_b200.circuit.Circuit` or a Qrisp ``QuantumCircuit``
(duck-typed through :meth:`Circuit.from_qrisp`).
"""

from __future__ import annotations

import numpy as np

from .circuit import Circuit
from .engine import StateVector
from .preprocess import count_measurements, strip_structural, unitarize

# reference: numerics_config.py:24-38
FLOAT_THRESH = np.float32(1e-5)
CUTOFF_RATIO = np.float32(2e-4)
# above this many measured qubits the dense distribution is not materialised on the host;
# shots are then sampled on the device straight from the amplitudes
MAX_DENSE_OUTCOME_BITS = 26
# host<->device bytes of the most recent run()/sample_counts()/statevector_sim() call
LAST_H2D_BYTES = 0
LAST_D2H_BYTES = 0


def _record_io(sv):
    global LAST_H2D_BYTES, LAST_D2H_BYTES
    LAST_H2D_BYTES, LAST_D2H_BYTES = sv.io["h2d"], sv.io["d2h"]


def _as_circuit(qc):
    return qc if isinstance(qc, Circuit) else Circuit.from_qrisp(qc)


def prune_outcomes(p):
    """Which outcomes the reference keeps, and their normalised probabilities.

    reference: bi_arrays.py:1030-1055 -- more than 10 measured qubits: keep
    ``p > max(p) * cutoff_ratio`` (bi_array_helper.py:315-331); otherwise the recursive
    halving of bi_array_helper.py:349-387 drops every branch whose norm falls below
    ``float_thresh``.  tensor_factor.py:184 / quantum_state.py:224 renormalise.
    """
    m = int(np.log2(p.size))
    if m > 10:
        keep = np.flatnonzero(p > p.max() * CUTOFF_RATIO)
    else:
        alive = np.ones(1, dtype=bool)
        for level in range(m + 1):
            branch = p.reshape(1 << level, -1).sum(axis=1)
            alive &= ~(branch < FLOAT_THRESH)
            if level < m:
                alive = np.repeat(alive, 2)
        keep = np.flatnonzero(alive)
    probs = p[keep]
    return keep.astype(np.int64), probs / probs.sum()


def round_probabilities(cl_prob):
    """reference: simulator.py:167-169."""
    cl_prob = np.round(cl_prob, int(-np.log10(np.max(cl_prob))) + 5)
    return cl_prob / np.sum(cl_prob)


def _device_outcomes(sv, mes_qubits):
    """Outcomes the reference keeps for more than 10 measured qubits, pruned ON THE DEVICE: the dense
    distribution (2**m doubles) stays in HBM, only ``p > max(p) * cutoff_ratio`` travels
    (reference: bi_arrays.py:1030-1055 -> bi_array_helper.py:315-331)."""
    import torch

    from . import _lib

    m = len(mes_qubits)
    probs = sv.probabilities(mes_qubits, to_host=False)
    scratch = torch.zeros(2, dtype=torch.int64, device=sv.device)
    ratio = float(CUTOFF_RATIO)
    stream = sv._stream()
    _lib.check(sv.lib.qb2_prune_probabilities(probs.data_ptr(), m, ratio, scratch.data_ptr(), None, None, 0, stream), "qb2_prune_probabilities")
    count = int(scratch[1].item())
    idx = torch.empty(count, dtype=torch.int64, device=sv.device)
    val = torch.empty(count, dtype=torch.float64, device=sv.device)
    _lib.check(
        sv.lib.qb2_prune_probabilities(probs.data_ptr(), m, ratio, scratch.data_ptr(), idx.data_ptr(), val.data_ptr(), count, stream),
        "qb2_prune_probabilities",
    )
    keep, p = idx.cpu().numpy(), val.cpu().numpy()
    sv.io["d2h"] += 16 * count + 16
    order = np.argsort(keep, kind="stable")  # the compaction's order is not deterministic; the reference's is ascending
    keep, p = keep[order], p[order]
    return keep.astype(np.int64), p / p.sum()


def run(qc, shots=None, token="", iqs=None, insert_reset=True, dtype=np.complex64, seed=None, return_engine=False,
        group=None, **engine_kwargs):
    """Simulate ``qc`` and return its measurement statistics -- the drop-in for
    ``qrisp.simulator.run(qc, shots, token="", iqs=None, insert_reset=True)`` (reference: simulator.py:44).

    ``shots=None``: dict ``bitstring -> probability``; ``shots=k``: dict ``bitstring -> count``.  ``token`` is
    accepted and ignored, as in the reference.  ``insert_reset`` only makes the reference sprinkle
    ``disentangle`` hints for its factored state (circuit_preprocessing.py:466-491): no effect on a dense
    state.  ``iqs`` (a reference ``QuantumState`` to start from) has no B200 counterpart.  ``seed`` fixes the
    sampled counts.  ``group``: a ``torch.distributed`` process group -- every rank calls ``run`` with the same
    arguments, the state is sharded over the ranks' GPUs and every rank gets the result.
    """
    if iqs is not None:
        raise NotImplementedError("run(iqs=...): starting from a reference QuantumState object is not supported on the B200 path")
    qc = _as_circuit(qc)
    if len(qc.data) == 0:
        return {"": shots}
    if shots == 0:
        return {}
    if count_measurements(qc.data) == 0:
        return {"": 1.0}
    instrs, n, measurements = unitarize(qc)
    if group is None:
        sv = StateVector(n, dtype=dtype, **engine_kwargs)
    else:
        from .distributed import ShardedStateVector

        sv = ShardedStateVector(n, group=group, dtype=dtype, **engine_kwargs)
    sv.apply_circuit(instrs)
    # reference: simulator.py:151-157 -- outcome bit order = descending clbit index
    measurements.sort(key=lambda x: -x[1])
    mes_qubits = [q for q, _ in measurements]
    width = len(mes_qubits)
    res = {}
    if width <= MAX_DENSE_OUTCOME_BITS:
        if width > 10 and group is None:
            outcomes, probs = _device_outcomes(sv, mes_qubits)
        else:
            outcomes, probs = prune_outcomes(sv.probabilities(mes_qubits))
        probs = round_probabilities(probs)
        if shots is None:
            for o, pr in zip(outcomes, probs):
                if pr == 0:
                    continue
                key = bin(int(o))[2:].zfill(width)
                res[key] = res.get(key, 0.0) + float(pr)
        else:
            rng = np.random.default_rng(seed)
            samples = rng.choice(len(probs), int(shots), p=probs)
            res = counts_dict(outcomes[samples].astype(np.uint64), width)
    else:
        if shots is None:
            raise MemoryError(
                f"{width} measured qubits: the full distribution has 2**{width} entries; pass shots=... to sample it"
            )
        # wide registers: shots are drawn on the device straight from the amplitudes
        outs = sv.sample(int(shots), mes_qubits, rng=np.random.default_rng(seed))
        res = counts_dict(outs, width)
    _record_io(sv)
    if return_engine:
        return res, sv
    return res


def counts_dict(outcomes, width):
    """{bitstring: count} of an array of outcome integers, most significant bit first (the
    reference's label convention, simulator.py:151-203).  The labels are built with numpy (bits ->
    ASCII bytes -> fixed-width strings): a Python loop over thousands of outcomes costs more than
    the device needs for a 30-qubit sweep pass."""
    vals, counts = np.unique(np.asarray(outcomes, dtype=np.uint64), return_counts=True)
    if width == 0:
        return {"": int(counts.sum())} if len(vals) else {}
    shifts = np.arange(width - 1, -1, -1, dtype=np.uint64)
    chars = (((vals[:, None] >> shifts[None, :]) & np.uint64(1)) + np.uint64(48)).astype(np.uint8)
    labels = np.ascontiguousarray(chars).view(f"S{width}").ravel()
    return dict(zip(labels.astype(f"U{width}").tolist(), counts.tolist()))


def sample_counts(qc, shots, seed=None, dtype=np.complex64, **engine_kwargs):
    """Counts sampled ON THE DEVICE from the amplitudes (no dense distribution on the host):
    the path for wide registers (``get_measurement`` of 30+ qubits)."""
    qc = _as_circuit(qc)
    instrs, n, measurements = unitarize(qc)
    sv = StateVector(n, dtype=dtype, **engine_kwargs)
    sv.apply_circuit(instrs)
    measurements.sort(key=lambda x: -x[1])
    mes_qubits = [q for q, _ in measurements]
    outs = sv.sample(int(shots), mes_qubits, rng=np.random.default_rng(seed))
    _record_io(sv)
    return counts_dict(outs, len(mes_qubits))


def statevector_sim(qc, dtype=np.complex64, **engine_kwargs):
    """The final statevector of a measurement-free circuit, qubit 0 most significant
    (reference: simulator.py:220-296)."""
    qc = _as_circuit(qc) if isinstance(qc, Circuit) else Circuit.from_qrisp(qc, transpile=False)
    if count_measurements(qc.data) != 0:
        raise Exception("Tried to determine the statevector of a circuit containing a measurement")
    instrs = strip_structural(qc.data)
    for i in instrs:
        if i.matrix is None:
            raise Exception(f"statevector_sim: non-unitary instruction {i.name!r}")
    sv = StateVector(qc.num_qubits, dtype=dtype, **engine_kwargs)
    sv.apply_circuit(instrs)
    res = sv.statevector()
    _record_io(sv)
    return res


from .backend import B200Backend, B200Job, B200JobResult, make_qrisp_backend  # noqa: E402,F401  (the default-backend drop-in)
