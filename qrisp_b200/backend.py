"""The default-backend drop-in: Qrisp's ``Backend`` contract on top of :func:`qrisp_b200.simulator.run`.

Reference: ``qrisp/interface/backend.py:83-330`` (``Backend``: ``run_async(circuits, shots) -> Job``,
``run(circuits, shots)``, runtime ``options``), ``qrisp/interface/job.py`` (``Job`` / ``JobResult`` /
``JobStatus``) and ``qrisp/interface/simulators/qrisp_simulator_backend.py:33-338``
(``QrispSimulatorBackend`` / ``QrispSimulatorJob``: synchronous, the job is DONE when ``run_async``
returns; default options ``{"shots": None, "token": ""}``), installed as
``qrisp.default_backend.def_backend``.

Two forms:

* :class:`B200Backend` has the same methods and semantics and needs no Qrisp install (tests, benches,
  hosts that hand over :class:`qrisp_b200.circuit.Circuit` objects);
* :func:`make_qrisp_backend` builds, inside a process that has Qrisp, a real subclass of
  ``qrisp.interface.Backend`` whose jobs are ``qrisp.interface.Job`` objects -- the object to assign to
  ``qrisp.default_backend.def_backend``.  Qrisp is imported only there.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence

import numpy as np


def _is_batch(circuits):
    return isinstance(circuits, Sequence) and not hasattr(circuits, "data")


def _validate_shots(shots):
    # reference: backend.py Backend._validate_shots
    if shots is None:
        return
    for s in shots if isinstance(shots, list) else [shots]:
        if isinstance(s, bool) or not isinstance(s, (int, np.integer)):
            raise TypeError(f"'shots' must be an integer, got {type(s).__name__}")
        if s <= 0:
            raise ValueError(f"'shots' must be a positive integer, got {s}")


class B200JobResult:
    """Counterpart of ``qrisp.interface.JobResult`` (job.py:83-187): one counts dict per circuit."""

    def __init__(self, counts, **kwargs):
        if not isinstance(counts, Sequence) or isinstance(counts, (str, bytes)):
            raise TypeError(f"'counts' must be a sequence of dicts, got {type(counts).__name__}")
        if len(counts) == 0:
            raise ValueError("'counts' must contain at least one dict")
        self._counts = list(counts)
        self.metadata = kwargs

    @property
    def all_counts(self):
        return self._counts

    @property
    def num_circuits(self):
        return len(self._counts)

    def get_counts(self, index=0):
        try:
            return self._counts[index]
        except IndexError as exc:
            raise IndexError(f"Result contains {len(self._counts)} circuit(s); index {index} is out of range.") from exc

    def __repr__(self):
        return f"B200JobResult(num_circuits={self.num_circuits}, metadata={self.metadata!r})"


class B200Job:
    """Counterpart of ``QrispSimulatorJob``: synchronous -- :meth:`submit` runs every circuit, so the job
    is ``"done"`` (or ``"error"``) by the time ``run_async`` hands it out."""

    def __init__(self, backend, circuits, shots):
        self._backend = backend
        self._circuits = circuits
        self._shots = shots
        self._result = None
        self._failure = None
        self._status = "initializing"

    def submit(self):
        self._status = "running"
        try:
            shots = self._shots if isinstance(self._shots, list) else [self._shots] * len(self._circuits)
            self._result = B200JobResult([self._backend._simulate(c, s) for c, s in zip(self._circuits, shots)])
            self._status = "done"
        except Exception as exc:  # the reference stores the cause and raises it from result()
            self._failure = exc
            self._status = "error"

    def status(self):
        return self._status

    def done(self):  # reference: job.py:383-405 -- done() is DONE only; the terminal states are three
        return self._status == "done"

    def running(self):
        return self._status == "running"

    def cancelled(self):
        return self._status == "cancelled"

    def in_final_state(self):
        return self._status in ("done", "error", "cancelled")

    def cancel(self):
        return False  # synchronous jobs cannot be cancelled after submission

    def result(self, timeout=None):
        if self._status == "error":
            raise RuntimeError(f"B200 simulation failed: {self._failure!r}") from self._failure
        return self._result


class B200Backend:
    """``backend.run(qc, shots)`` / ``backend.run_async(circuits, shots).result().all_counts`` on the B200
    statevector engine.  ``circuits``: one circuit or a sequence (``QuantumCircuit`` or
    :class:`qrisp_b200.circuit.Circuit`); ``shots``: int, list of ints or None (the ``shots`` option;
    None = exact probabilities, as with the reference's simulator backend)."""

    def __init__(self, name=None, options=None, dtype=np.complex64, pm=None, **engine_kwargs):
        self.name = name or "B200Simulator"
        # reference: QrispSimulatorBackend(pm=...) -- a pass manager every circuit goes through before it is
        # simulated; anything with a ``run(circuit) -> circuit`` method (a qrisp PassManager has one)
        if pm is not None and not callable(getattr(pm, "run", None)):
            raise TypeError(f"Expected a PassManager instance for 'pm', got {type(pm).__name__}.")
        self._pm = pm
        if options is None:
            options = self._default_options()
        if not isinstance(options, Mapping):
            raise TypeError(f"'options' must be a dict-like Mapping, got {type(options).__name__}")
        self._options = dict(options)
        self.dtype = dtype
        self.engine_kwargs = engine_kwargs

    @classmethod
    def _default_options(cls):
        return {"shots": None, "token": ""}

    @property
    def options(self):
        return self._options

    def update_options(self, **kwargs):
        for k, v in kwargs.items():  # reference: only keys present at initialisation may be modified
            if k not in self._options:
                raise AttributeError(f"'{k}' is not a valid option for backend '{self.name}'")
            self._options[k] = v

    def _simulate(self, circuit, shots):
        from .simulator import run

        return run(circuit, shots, self._options.get("token", ""), dtype=self.dtype, **self.engine_kwargs)

    def run_async(self, circuits, shots=None):
        _validate_shots(shots)
        batch = list(circuits) if _is_batch(circuits) else [circuits]
        if self._pm is not None:
            batch = [self._pm.run(c) for c in batch]
        if isinstance(shots, list) and len(shots) != len(batch):
            raise ValueError(f"Length of 'shots' ({len(shots)}) must match the number of circuits ({len(batch)}).")
        if shots is None:
            shots = self._options.get("shots")
        job = B200Job(self, batch, shots)
        job.submit()
        return job

    def run(self, circuits, shots=None, token=None):
        """One counts / probability dict for a single circuit, a list of them for a sequence (the reference
        wraps them in ``MeasurementResult``, a lazily decoded dict: backend.py:284-330)."""
        if token is not None:
            self._options["token"] = token
        res = self.run_async(circuits, shots).result().all_counts
        return res if _is_batch(circuits) else res[0]


def make_qrisp_backend(dtype=np.complex64, pm=None, **engine_kwargs):
    """A ``qrisp.interface.Backend`` that simulates on the B200: assign it to
    ``qrisp.default_backend.def_backend`` (or pass ``backend=...`` to ``get_measurement``).  ``pm``: a Qrisp
    ``PassManager`` applied to every circuit first, as ``QrispSimulatorBackend(pm=...)`` does
    (qrisp_simulator_backend.py:262-330).  Needs Qrisp importable in the calling process; nothing else in this
    package does."""
    from qrisp.circuit.pass_management import PassManager
    from qrisp.interface.backend import Backend
    from qrisp.interface.job import Job, JobResult, JobStatus

    from .simulator import run as b200_run

    class B200QrispJob(Job):
        def __init__(self, backend, circuits, shots):
            super().__init__(backend=backend)
            self._circuits, self._shots, self._result_data = circuits, shots, None

        def submit(self):
            token = self._backend.options.get("token", "")
            self._last_known_status = JobStatus.RUNNING
            try:
                shots = self._shots if isinstance(self._shots, list) else [self._shots] * len(self._circuits)
                self._result_data = JobResult(
                    [b200_run(c, s, token, dtype=dtype, **engine_kwargs) for c, s in zip(self._circuits, shots)]
                )
                self._last_known_status = JobStatus.DONE
            except Exception as exc:
                self._failure_cause = exc
                self._last_known_status = JobStatus.ERROR

        def result(self, timeout=None):
            self._raise_for_status(self._last_known_status)
            return self._result_data

        def cancel(self):
            return False

        def status(self):
            return self._last_known_status

    class B200QrispBackend(Backend):
        def __init__(self, pm=None):
            super().__init__(name="B200Simulator", options=None)
            if pm is not None and not isinstance(pm, PassManager):
                raise TypeError(f"Expected a PassManager instance for 'pm', got {type(pm).__name__}.")
            self._pm = pm

        @classmethod
        def _default_options(cls):
            return {"shots": None, "token": ""}

        def run_async(self, circuits, shots=None):
            self._check_circuit_limit(circuits)
            batch = list(circuits) if _is_batch(circuits) else [circuits]
            if self._pm is not None:
                batch = [self._pm.run(c) for c in batch]
            if isinstance(shots, list):
                self._validate_shots_length(shots, batch)
            if shots is None:
                shots = self.options.get("shots")
            job = B200QrispJob(self, batch, shots)
            job.submit()
            return job

    return B200QrispBackend(pm)
