"""The Backend contract without a GPU (reference tests: tests/interface_tests/test_default_backend.py):
options, the ``pm`` keyword, argument validation and the failure path -- everything in front of the first launch.
The Qrisp form is checked against the unmodified reference package (oracle/_ref) when it is built."""

import pytest

from qrisp_b200.backend import B200Backend, B200JobResult, make_qrisp_backend
from qrisp_b200.circuit import Circuit


def _ref():
    from oracle import refstub

    qrisp = refstub.import_reference()
    if qrisp is None:
        pytest.skip("oracle/_ref is not built (python oracle/build_ref.py, in the build container)")
    return qrisp


def test_options_follow_the_reference_backend():
    b1, b2 = B200Backend(), B200Backend()
    assert b1.options["shots"] is None and b1.options["token"] == ""
    b1.update_options(shots=512)
    assert b1.options["shots"] == 512 and b2.options["shots"] is None  # instances do not share options
    with pytest.raises(AttributeError):
        b1.update_options(nonexistent=42)
    with pytest.raises(TypeError):
        B200Backend(options=3)


def test_shots_and_pm_are_validated_before_anything_runs():
    b = B200Backend()
    qc = Circuit(2)
    for bad in (0, -3):
        with pytest.raises(ValueError):
            b.run_async(qc, shots=bad)
    for bad in (True, 1.5, "7"):
        with pytest.raises(TypeError):
            b.run_async(qc, shots=bad)
    with pytest.raises(ValueError):
        b.run_async([qc, qc], shots=[10])  # one shot count per circuit
    with pytest.raises(TypeError):
        B200Backend(pm="not_a_pass_manager")


def test_pm_runs_on_every_circuit_and_a_failure_surfaces_from_result(monkeypatch):
    """No device here: ``run`` fails inside the job, which must end in "error", keep the cause and raise it
    from ``result()`` -- after the pass manager has seen every circuit."""
    seen = []

    class Passes:
        def run(self, qc):
            seen.append(qc)
            return qc

    b = B200Backend(pm=Passes())

    def boom(circuit, shots):
        raise KeyError("unknown_gate")

    monkeypatch.setattr(b, "_simulate", boom)
    qc = Circuit(1)
    job = b.run_async([qc, qc, qc], shots=5)
    assert len(seen) == 3
    assert job.status() == "error" and job.in_final_state() and not job.done()
    assert job.cancel() is False and job.status() == "error"
    with pytest.raises(RuntimeError, match="unknown_gate") as info:
        job.result()
    assert isinstance(info.value.__cause__, KeyError)


def test_job_result_contract():
    r = B200JobResult([{"0": 1.0}, {"1": 1.0}], note="x")
    assert r.num_circuits == 2 and r.get_counts() == {"0": 1.0} and r.get_counts(1) == {"1": 1.0} and r.metadata == {"note": "x"}
    with pytest.raises(IndexError):
        r.get_counts(2)
    with pytest.raises(ValueError):
        B200JobResult([])
    with pytest.raises(TypeError):
        B200JobResult("01")


def test_qrisp_form_is_a_backend_of_the_reference_package():
    """``make_qrisp_backend()`` against the unmodified reference: a ``qrisp.interface.Backend`` with the
    simulator backend's options and ``pm`` handling; a circuit the simulator cannot run gives a job in ERROR whose
    ``result()`` raises ``JobFailureError`` naming the gate, with the cause chained (the failure happens while the
    circuit is read, in front of any launch, so this runs without a GPU)."""
    _ref()
    from qrisp import QuantumCircuit
    from qrisp.circuit import Operation
    from qrisp.circuit.pass_management import PassManager
    from qrisp.interface import Backend
    from qrisp.interface.job import Job, JobFailureError, JobStatus

    backend = make_qrisp_backend()
    assert isinstance(backend, Backend)
    assert backend.options["shots"] is None and backend.options["token"] == ""
    backend.update_options(shots=512)
    assert backend.options["shots"] == 512 and make_qrisp_backend().options["shots"] is None
    with pytest.raises(AttributeError):
        backend.update_options(nonexistent=42)
    with pytest.raises(TypeError):
        make_qrisp_backend(pm="not_a_pass_manager")
    assert make_qrisp_backend(pm=PassManager())._pm is not None

    qc = QuantumCircuit(1, 1)
    qc.append(Operation(name="unknown_gate", num_qubits=1), [0])
    qc.measure(0, 0)
    job = make_qrisp_backend().run_async(qc)
    assert isinstance(job, Job) and job.status() == JobStatus.ERROR and job.in_final_state() and not job.done()
    assert job.cancel() is False
    with pytest.raises(JobFailureError, match="unknown_gate") as info:
        job.result()
    assert info.value.__cause__ is not None
    with pytest.raises(ValueError):
        make_qrisp_backend().run_async([qc, qc], shots=[1, 2, 3])
