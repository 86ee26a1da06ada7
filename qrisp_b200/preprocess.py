"""Circuit rewriting that precedes the statevector sweep (host logic, no amplitudes).

Mirrors what ``qrisp.simulator.run`` does before its main loop:

* structural instructions (barrier / qb_alloc / qb_dealloc) are dropped
  (reference: circuit_preprocessing.py:466-491 ``count_measurements_and_treat_alloc``);
* every mid-circuit measurement and every reset is replaced by CX gates onto a fresh qubit,
  so that the whole circuit becomes one unitary followed by terminal measurements
  (reference: circuit_preprocessing.py:729-836 ``insert_multiverse_measurements``).

The reference also sprinkles ``disentangle`` hints into the stream (they let its lazily
factored ``TensorFactor`` state split off qubits, tensor_factor.py:198-247).  They never
change the state, and the B200 engine keeps one dense statevector, so they are dropped.
"""

from __future__ import annotations

import numpy as np

from .circuit import STRUCTURAL, Instruction, _controlled

_CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)


def commutes_with_z(instr, qubit, tol=1e-10):
    """True when ``instr`` is block diagonal in ``qubit`` ("permeable" in the reference's
    terms: permeability/type_checker.py:48)."""
    if instr.matrix is None:
        return False
    k = len(instr.qubits)
    bit = k - 1 - instr.qubits.index(qubit)  # big-endian matrix index
    rows, cols = np.nonzero(np.abs(instr.matrix) > tol)
    return not np.any(((rows ^ cols) >> bit) & 1)


def strip_structural(data):
    return [i for i in data if i.name not in STRUCTURAL and i.name != "disentangle"]


def count_measurements(data):
    return sum(1 for i in data if i.name == "measure")


def unitarize(circuit):
    """Return ``(unitary_instructions, num_qubits, measurements)``.

    ``measurements`` is a list of ``(qubit, clbit)`` terminal measurements; ``num_qubits``
    includes the fresh qubits that stand in for mid-circuit measurement outcomes and
    discarded (reset) branches.
    """
    n = circuit.num_qubits
    stream = strip_structural(circuit.data)
    unitary = []
    terminal = []
    outcome_qubit = {}  # clbit -> fresh qubit that carries the outcome
    idx = 0
    while idx < len(stream):
        instr = stream[idx]
        idx += 1
        if instr.name == "measure":
            q, c = instr.qubits[0], instr.clbits[0]
            reset_follows = False
            is_terminal = True
            look = idx
            while look < len(stream):
                nxt = stream[look]
                if q in nxt.qubits:
                    if nxt.name == "reset":
                        # NOTE (pinned by tests/test_planner_cpu.py::test_reset_is_hoisted_over_z_commuting_gates):
                        # the reset is taken out of the stream HERE, i.e. it moves in front of the gates in
                        # between, all of which commute with Z on q.  A gate that q only CONTROLS (measure q;
                        # cx(q, r); reset q) therefore sees q after the reset and never fires -- exactly what the
                        # reference's insert_multiverse_measurements does (circuit_preprocessing.py:755-789:
                        # the same look-ahead, `reset_follows` + `instructions.pop`); kept for parity.
                        reset_follows = True
                        del stream[look]
                        is_terminal = False
                        break
                    if not commutes_with_z(nxt, q):
                        is_terminal = False
                        break
                if c in nxt.clbits:
                    is_terminal = False
                    break
                look += 1
            if is_terminal:
                terminal.append((q, c))
                continue
            fresh = n
            n += 1
            unitary.append(Instruction("cx", (q, fresh), (), _CX))
            if reset_follows:
                unitary.append(Instruction("cx", (fresh, q), (), _CX))
            outcome_qubit[c] = fresh
        elif instr.name == "reset":
            q = instr.qubits[0]
            needed = any(q in nxt.qubits and not commutes_with_z(nxt, q) for nxt in stream[idx:])
            if not needed:
                continue
            fresh = n
            n += 1
            unitary.append(Instruction("cx", (q, fresh), (), _CX))
            unitary.append(Instruction("cx", (fresh, q), (), _CX))
        elif instr.ctrl_state is not None:
            # classically controlled operation (reference: circuit_preprocessing.py:790-816): a control on
            # the qubit that carries the clbit's outcome.  A clbit nobody has written yet reads 0: the
            # operation never fires if it wants a 1, and a fresh |0> qubit stands in if it wants a 0.
            controls = []
            fires = True
            for c, want in zip(instr.clbits, instr.ctrl_state):
                if c in outcome_qubit:
                    controls.append(outcome_qubit[c])
                elif want == "1":
                    fires = False
                    break
                else:
                    controls.append(n)
                    n += 1
            if fires:
                m = _controlled(instr.matrix, len(controls), int(instr.ctrl_state, 2)) if controls else instr.matrix
                unitary.append(Instruction(instr.name, tuple(controls) + instr.qubits, (), m))
        elif instr.matrix is None:
            raise NotImplementedError(f"instruction {instr.name!r} has no unitary and is not measure/reset")
        else:
            unitary.append(instr)
    for c, q in outcome_qubit.items():
        terminal.append((q, c))
    terminal = list(dict.fromkeys(terminal))
    return unitary, n, terminal
