"""qrisp_b200 -- a B200-native statevector engine behind Qrisp's simulator interface.

Drop-in scope: the path behind ``qrisp.simulator.run(qc, shots, token)`` /
``statevector_sim(qc)`` and the default backend.  Importing the package does
not need a GPU; creating a :class:`StateVector` does, and fails loudly without one.
"""

from .circuit import Circuit, Instruction, ghz_circuit, ghz_qft_circuit, qft_circuit, random_circuit  # noqa: F401

__all__ = [
    "Circuit",
    "Instruction",
    "ghz_circuit",
    "ghz_qft_circuit",
    "qft_circuit",
    "random_circuit",
    "run",
    "statevector_sim",
    "StateVector",
    "B200Backend",
]


def __getattr__(name):
    if name in ("run", "statevector_sim", "B200Backend", "sample_counts"):
        from . import simulator

        return getattr(simulator, name)
    if name == "StateVector":
        from .engine import StateVector

        return StateVector
    raise AttributeError(name)
