"""Pins the CPU oracle (oracle/ref_sim.py) and the host gate library against fixtures that
were produced by running the reference itself (tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

from oracle import ref_sim
from qrisp_b200 import circuit as C
from qrisp_b200.circuit import Circuit


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_gate_library_matches_reference_unitaries(golden_dir):
    g = load(golden_dir, "gates.npz")
    build = {
        "h": lambda qc: qc.h(0), "x": lambda qc: qc.x(0), "y": lambda qc: qc.y(0), "z": lambda qc: qc.z(0),
        "s": lambda qc: qc.s(0), "s_dg": lambda qc: qc.s_dg(0), "t": lambda qc: qc.t(0),
        "t_dg": lambda qc: qc.t_dg(0), "sx": lambda qc: qc.sx(0), "p": lambda qc: qc.p(0.37, 0),
        "rx": lambda qc: qc.rx(0.37, 0), "ry": lambda qc: qc.ry(0.37, 0), "rz": lambda qc: qc.rz(0.37, 0),
        "u3": lambda qc: qc.u3(0.37, 1.1, -0.6, 0), "gphase": lambda qc: qc.gphase(0.37, 0),
        "cx": lambda qc: qc.cx(0, 1), "cy": lambda qc: qc.cy(0, 1), "cz": lambda qc: qc.cz(0, 1),
        "cp": lambda qc: qc.cp(0.37, 0, 1), "crz": lambda qc: qc.crz(0.37, 0, 1),
        "swap": lambda qc: qc.swap(0, 1), "rzz": lambda qc: qc.rzz(0.37, 0, 1),
        "rxx": lambda qc: qc.rxx(0.37, 0, 1), "mcx": lambda qc: qc.mcx([0, 1], 2),
        "mcp": lambda qc: qc.mcp(0.37, [0, 1], 2),
        "sx_dg": lambda qc: qc.sx_dg(0), "id": lambda qc: qc.id(0), "r": lambda qc: qc.r(0.37, 1.1, 0),
        "crx": lambda qc: qc.crx(0.37, 0, 1), "xxyy": lambda qc: qc.xxyy(0.37, 1.1, 0, 1),
    }
    assert set(build) == set(str(n) for n in g["names"])
    for name in g["names"]:
        name = str(name)
        ref = g["op_" + name]
        k = int(np.log2(ref.shape[0]))
        qc = Circuit(k)
        build[name](qc)
        assert tuple(g["opq_" + name]) == qc.data[0].qubits, name
        # the reference computes its matrices in complex64
        np.testing.assert_allclose(qc.data[0].matrix, ref, atol=2e-7, err_msg=name)


@pytest.mark.parametrize("threshold", [False, True])
def test_oracle_statevector_matches_reference(golden_dir, threshold):
    g = load(golden_dir, "statevector.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        sv = ref_sim.statevector(qc, dtype=np.complex128, threshold=threshold)
        # north_star tolerance: 1e-5 absolute at complex64
        np.testing.assert_allclose(sv, g[case + "_sv"], atol=1e-5, err_msg=case)


def test_oracle_reference_path_matches_reference(golden_dir):
    """The CPU arm of bench.py (grouped unitaries, sparse while the state is sparse) reaches
    the reference's statevectors too."""
    g = load(golden_dir, "statevector.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        for kmax in (3, 6):
            sv = ref_sim.statevector_reference_path(qc, dtype=np.complex128, kmax=kmax)
            np.testing.assert_allclose(sv, g[case + "_sv"], atol=1e-5, err_msg=case)
    # a state that stays sparse to the end (GHZ) and one that turns dense on the way
    from qrisp_b200.circuit import ghz_qft_circuit

    qc = ghz_qft_circuit(11)
    np.testing.assert_allclose(ref_sim.statevector_reference_path(qc, dtype=np.complex128),
                               ref_sim.statevector(qc, dtype=np.complex128), atol=1e-10)


def _as_dict(keys, probs):
    return {str(k): float(p) for k, p in zip(keys, probs)}


def _assert_dist_close(got, ref, atol=1e-5):
    for k in set(got) | set(ref):
        assert abs(got.get(k, 0.0) - ref.get(k, 0.0)) <= atol, (k, got.get(k), ref.get(k))


@pytest.mark.parametrize("fixture", ["run.npz", "programs.npz", "cif.npz", "algorithms.npz"])
def test_oracle_run_matches_reference(golden_dir, fixture):
    g = load(golden_dir, fixture)
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        if qc.num_qubits > 24:
            continue  # too wide for the dense numpy oracle; covered by the factored host path
        got = ref_sim.run(qc, None, dtype=np.complex128)
        ref = _as_dict(g[case + "_keys"], g[case + "_probs"])
        # basis-state labels must agree exactly for every outcome above the tolerance
        assert {k for k, v in ref.items() if v > 1e-4} <= set(got), case
        _assert_dist_close(got, ref)


def test_oracle_circuit_unitary_matches_reference_groups(golden_dir):
    g = load(golden_dir, "fusion.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        goff, gq, gm = g[case + "_goff"], g[case + "_gq"], g[case + "_gm"]
        # replaying the reference's grouped unitaries through the oracle must give the
        # reference's statevector: pins apply_matrix for k-qubit fused matrices
        psi = ref_sim.zero_state(qc.num_qubits, np.complex128)
        moff = 0
        for j in range(len(goff) - 1):
            qubits = gq[goff[j] : goff[j + 1]]
            dim = 1 << len(qubits)
            psi = ref_sim.apply_matrix(psi, gm[moff : moff + dim * dim].reshape(dim, dim), list(qubits))
            moff += dim * dim
        np.testing.assert_allclose(psi, g[case + "_sv"], atol=1e-5, err_msg=case)
        # and the whole circuit as one fused unitary does too: pins circuit_unitary
        u = ref_sim.circuit_unitary(qc.data, list(range(qc.num_qubits)))
        np.testing.assert_allclose(u[:, 0], g[case + "_sv"], atol=1e-5, err_msg=case)


def test_oracle_edge_cases():
    qc = Circuit(2)
    assert ref_sim.run(qc, 100) == {"": 100}          # reference: simulator.py:46-47
    qc.h(0)
    assert ref_sim.run(qc, 0) == {}                   # reference: simulator.py:48-49
    assert ref_sim.run(qc, None) == {"": 1.0}         # reference: simulator.py:74-77
    qc.measure([0, 1])
    with pytest.raises(Exception):
        ref_sim.statevector(qc)                       # reference: simulator.py:251-252
    counts = ref_sim.run(qc, 1000, seed=1)
    assert sum(counts.values()) == 1000 and set(counts) <= {"00", "01"}
