"""GPU parity tests: the CUDA path (through the C-ABI of libqb2.so) against the oracle, the
numpy pass emulator and the golden fixtures produced by the reference itself.

Tolerances (north_star): amplitudes and probabilities within 1e-5 absolute at complex64,
1e-12 at complex128; basis-state labels and qubit order bit-exact; sampled counts consistent
by a chi-square test.
"""

import os

import numpy as np
import pytest

from oracle import pass_emulator, ref_sim
from qrisp_b200.circuit import Circuit, ghz_circuit, ghz_qft_circuit, qft_circuit, random_circuit

from helpers import plan_circuit

pytestmark = pytest.mark.gpu

ATOL64 = 1e-5
ATOL128 = 1e-12


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.fixture(scope="module")
def sim():
    from qrisp_b200 import simulator

    return simulator


# ---------------------------------------------------------------------------------------
# statevectors
# ---------------------------------------------------------------------------------------


def test_golden_statevectors_c64(golden_dir, sim):
    g = load(golden_dir, "statevector.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        sv = sim.statevector_sim(qc)
        assert sv.dtype == np.complex64 and sv.shape == (1 << qc.num_qubits,)
        np.testing.assert_allclose(sv, g[case + "_sv"], atol=ATOL64, err_msg=case)


def test_golden_statevectors_c128(golden_dir, sim):
    g = load(golden_dir, "statevector.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        sv = sim.statevector_sim(qc, dtype=np.complex128)
        want = ref_sim.statevector(qc, dtype=np.complex128)
        np.testing.assert_allclose(sv, want, atol=ATOL128, err_msg=case)


def test_golden_fusion_circuits(golden_dir, sim):
    g = load(golden_dir, "fusion.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        np.testing.assert_allclose(sim.statevector_sim(qc), g[case + "_sv"], atol=ATOL64, err_msg=case)


@pytest.mark.parametrize("T,r,L", [(9, 4, 4), (10, 3, 4), (11, 4, 5), (12, 4, 4), (13, 5, 4), (13, 4, 4), (12, 5, 6)])
def test_tile_geometries(sim, T, r, L):
    qc = random_circuit(15, 5, seed=T * 10 + r, two_qubit="cx")
    qc.rxx(0.3, 0, 14)
    qc.mcx([1, 5, 9], 0)
    qc.swap(2, 11)
    qc.h(11)
    sv = sim.statevector_sim(qc, T=T, r=r, L=L)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc), atol=ATOL64)


@pytest.mark.parametrize("dtype,r", [(np.complex128, 3), (np.complex128, 4)])
def test_c128_geometries(sim, dtype, r):
    qc = random_circuit(14, 5, seed=r, two_qubit="cx")
    qc.mcp(0.4, [0, 3, 13], 7)
    sv = sim.statevector_sim(qc, dtype=dtype, r=r, T=r + 8)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc, dtype=np.complex128), atol=ATOL128)


def test_gpu_matches_emulator_bitwise_layout():
    """Every packed pass gives the same amplitudes on the GPU as in the numpy emulator
    (different arithmetic order, so a tolerance; same layout, so no permutation slack)."""
    import torch

    from qrisp_b200.engine import StateVector

    qc = random_circuit(14, 6, seed=77, two_qubit="cx")
    qc.cp(0.3, 0, 13)
    qc.mcx([0, 13], 5)
    sv = StateVector(14, prefix_support=0)  # no host prefix: the same gate list goes to both planners
    instrs = [i for i in qc.data if i.matrix is not None]
    steps, pos, N = plan_circuit(instrs, 14)
    prog = sv.compile(instrs)
    assert [len(a.blob) for a in prog.steps] == [len(b.blob) for b in steps]
    st = np.zeros(1 << 14, dtype=np.complex64)
    st[0] = 1
    for a, b in zip(prog.steps, steps):
        assert a.blob == b.blob  # planning is deterministic
        st = pass_emulator.run(b.blob, st)
    prog.run()
    got = sv.state.cpu().numpy().view(np.complex64)
    np.testing.assert_allclose(got, st, atol=2e-6)


def test_dense_3q_4q_and_big_matrix(sim):
    rng = np.random.default_rng(5)

    def haar(k):
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _r = np.linalg.qr(a)
        return q

    qc = Circuit(13)
    for q in range(13):
        qc.h(q)
    qc.unitary(haar(3), (7, 0, 4))
    qc.unitary(haar(4), (10, 2, 9, 5))
    qc.unitary(haar(3), (1, 2, 3))
    qc.unitary(haar(5), (0, 3, 6, 8, 12))
    qc.unitary(haar(7), (11, 1, 5, 2, 9, 0, 4))
    qc.cz(0, 1)
    want = ref_sim.statevector(qc)
    np.testing.assert_allclose(sim.statevector_sim(qc), want, atol=ATOL64)
    np.testing.assert_allclose(sim.statevector_sim(qc, T=13, r=5), want, atol=ATOL64)
    want128 = ref_sim.statevector(qc, dtype=np.complex128)
    np.testing.assert_allclose(sim.statevector_sim(qc, dtype=np.complex128), want128, atol=ATOL128)


@pytest.mark.parametrize("n", [1, 2, 5, 9, 10])
def test_small_registers(sim, n):
    qc = ghz_qft_circuit(n)
    np.testing.assert_allclose(sim.statevector_sim(qc), ref_sim.statevector(qc), atol=ATOL64)
    empty = Circuit(n)
    sv = sim.statevector_sim(empty)
    assert sv[0] == 1 and np.count_nonzero(sv) == 1


def test_medium_width_against_oracle(sim):
    """22 qubits: still seconds for the numpy oracle, already 2**10 tiles for the kernel."""
    qc = random_circuit(22, 6, seed=1)
    np.testing.assert_allclose(sim.statevector_sim(qc), ref_sim.statevector(qc), atol=ATOL64)
    qc = ghz_qft_circuit(22)
    np.testing.assert_allclose(sim.statevector_sim(qc), ref_sim.statevector(qc), atol=ATOL64)


# ---------------------------------------------------------------------------------------
# run(): measurement statistics against what the reference returned
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("fixture", ["run.npz", "programs.npz", "cif.npz"])
def test_run_matches_reference_results(golden_dir, sim, fixture):
    g = load(golden_dir, fixture)
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        got = sim.run(qc, None)
        ref = {str(k): float(p) for k, p in zip(g[case + "_keys"], g[case + "_probs"])}
        assert set(got) == set(ref), case  # basis-state labels bit-exact
        for k in ref:
            assert abs(got[k] - ref[k]) < ATOL64, (case, k)


def test_run_edge_cases(sim):
    assert sim.run(Circuit(3), 100) == {"": 100}
    qc = Circuit(2)
    qc.h(0)
    assert sim.run(qc, None) == {"": 1.0}
    qc.measure([0, 1])
    assert sim.run(qc, 0) == {}
    res = sim.run(qc, 1000, seed=1)
    assert set(res) <= {"00", "01"} and sum(res.values()) == 1000
    with pytest.raises(Exception):
        sim.statevector_sim(qc)


def test_probabilities_and_norm():
    from qrisp_b200.engine import StateVector

    qc = random_circuit(16, 5, seed=9)
    psi = ref_sim.statevector(qc, dtype=np.complex128)
    for dtype, atol in ((np.complex64, 1e-6), (np.complex128, 1e-13)):
        sv = StateVector(16, dtype=dtype)
        sv.apply_circuit([i for i in qc.data if i.matrix is not None])
        assert abs(sv.norm2() - 1.0) < 10 * atol
        for qubits in ([0], [15], [3, 1, 12], list(range(16)), [15, 0, 7, 8, 2, 5, 11, 13, 1, 14, 3, 9], []):
            want = ref_sim.marginal_probabilities(psi, qubits, 16)
            got = sv.probabilities(qubits)
            np.testing.assert_allclose(got, want, atol=atol, err_msg=str(qubits))


def test_sampling_chi_square():
    """Counts drawn on the device are consistent with the oracle's distribution."""
    from qrisp_b200.engine import StateVector

    qc = random_circuit(12, 4, seed=4)
    sv = StateVector(12)
    sv.apply_circuit([i for i in qc.data if i.matrix is not None])
    psi = ref_sim.statevector(qc, dtype=np.complex128)
    qubits = [0, 5, 11, 3, 7]
    p = ref_sim.marginal_probabilities(psi, qubits, 12)
    shots = 200000
    outs = sv.sample(shots, qubits, rng=np.random.default_rng(123))
    counts = np.bincount(outs.astype(np.int64), minlength=32)
    chi2 = float(((counts - shots * p) ** 2 / (shots * p)).sum())
    # 31 degrees of freedom: the 99.99 % quantile is ~66
    assert chi2 < 70, chi2
    # the same seed gives the same counts
    outs2 = sv.sample(shots, qubits, rng=np.random.default_rng(123))
    assert np.array_equal(outs, outs2)
    # a basis state is sampled with certainty, with the right label
    sv2 = StateVector(11)
    c = Circuit(11)
    c.x(0), c.x(9)
    sv2.apply_circuit(c.data)
    outs = sv2.sample(100, list(range(11)), rng=np.random.default_rng(0))
    assert np.all(outs == int("10000000010", 2))


def test_replayed_program_can_leave_the_sampler_sums():
    """``Program.want_norms``: the last pass also writes the per-tile sums the sampler starts from; the draws
    are those of the separate reduction (same seed)."""
    from qrisp_b200.engine import StateVector

    n = 18
    qc = random_circuit(n, 2, seed=8)
    rng = np.random.default_rng(1)
    for first in (0, 1):  # two brick layers on qubits 0..12: the plan ends with a pass over the lowest 13 positions
        for q in range(first, 12, 2):
            a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            qc.unitary(np.linalg.qr(a)[0], [q, q + 1])
    instrs = [i for i in qc.data if i.matrix is not None]
    draws = []
    for want in (False, True):
        sv = StateVector(n)
        prog = sv.compile(instrs)
        prog.want_norms = want
        sv.reset()
        prog.run()
        if not want:
            assert sv._norms is None
        draws.append(sv.sample(5000, list(range(n)), rng=np.random.default_rng(5)))
        used = sv.stats.get("fused_norms", 0)
    assert used >= 1, "the sampler did not start from the pass's sums"
    # the two reductions round differently (per tile vs per chunk), so a uniform that falls within ~1e-7 of a
    # boundary may land on the neighbour
    assert np.mean(draws[0] == draws[1]) > 0.999


def test_wide_sampling_path(sim):
    """More measured qubits than the dense-distribution limit: counts come from the
    device sampler (GHZ: only all-zeros / all-ones)."""
    n = 27
    qc = ghz_circuit(n)
    qc.measure(list(range(n)))
    res = sim.run(qc, 2000, seed=3)
    assert set(res) == {"0" * n, "1" * n}
    assert sum(res.values()) == 2000 and abs(res["0" * n] - 1000) < 150


# ---------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: 30 qubits, complex64, one B200)
# ---------------------------------------------------------------------------------------


def test_full_size_30_qubits_properties():
    import torch

    from qrisp_b200.engine import StateVector

    n = 30
    sv = StateVector(n)
    # GHZ: exactly two amplitudes of 1/sqrt(2)
    sv.apply_circuit(ghz_circuit(n).data)
    assert abs(sv.norm2() - 1.0) < 1e-6
    st = sv.state.view(-1, 2)
    assert abs(float(st[0, 0]) - 2**-0.5) < 1e-6 and abs(float(st[-1, 0]) - 2**-0.5) < 1e-6
    assert int(torch.count_nonzero(st)) == 2
    # QFT then inverse QFT is the identity (round trip through ~1000 gates)
    sv.apply_circuit(qft_circuit(n).data)
    assert abs(sv.norm2() - 1.0) < 1e-5
    p = sv.probabilities([0, 1, 2])
    # QFT of GHZ: even outcomes only -> the least significant QFT output bit is 0;
    sv.apply_circuit(qft_circuit(n, inverse=True).data)
    st = sv.state.view(-1, 2)
    assert abs(float(st[0, 0]) - 2**-0.5) < 1e-4 and abs(float(st[-1, 0]) - 2**-0.5) < 1e-4
    rest = float(sv.norm2()) - float(st[0, 0]) ** 2 - float(st[-1, 0]) ** 2 - float(st[0, 1]) ** 2 - float(st[-1, 1]) ** 2
    assert abs(rest) < 1e-4
    assert abs(p.sum() - 1.0) < 1e-6
    # sampling labels at full width
    outs = sv.sample(64, list(range(n)), rng=np.random.default_rng(1))
    assert set(int(o) for o in outs) <= {0, (1 << n) - 1}


def ghz_qft_closed_form(n, y):
    """<y| QFT (|0..0> + |1..1>)/sqrt(2): 2**(-(n+1)/2) * (1 + exp(2 pi i y (2**n - 1) / 2**n)) for the textbook
    QFT of qrisp_b200.circuit.qft_circuit (final swaps included; the index has qubit 0 as its most
    significant bit).  Checked against the oracle in the test below before it is trusted."""
    y = np.asarray(y, dtype=np.float64)
    return 2.0 ** (-(n + 1) / 2) * (1.0 + np.exp(-2j * np.pi * y / 2.0**n))


def test_closed_form_of_ghz_qft_against_the_oracle():
    n = 10
    want = ref_sim.statevector(ghz_qft_circuit(n), dtype=np.complex128)
    np.testing.assert_allclose(ghz_qft_closed_form(n, np.arange(1 << n)), want, atol=1e-12)


@pytest.mark.parametrize("n", [30, 33])
def test_full_size_ghz_qft_matches_the_closed_form(n):
    """BASELINE.json configs[1] (30 qubits) and the widest single-GPU register (33 qubits, 64 GiB): amplitudes
    at random basis states against the closed form, within 1e-5 RELATIVE to the amplitude scale 2**(-n/2)
    (the absolute north-star tolerance would be vacuous at this width), and a chi-square of device
    samples against the closed-form distribution, labels bit-exact."""
    import torch

    from qrisp_b200.engine import StateVector

    free, _total = torch.cuda.mem_get_info()
    if free < (8 << n) + (2 << 30):
        pytest.skip(f"needs {(8 << n) >> 30} GiB of device memory")
    sv = StateVector(n)
    sv.apply_circuit([i for i in ghz_qft_circuit(n).data if i.matrix is not None])
    assert abs(sv.norm2() - 1.0) < 1e-5
    rng = np.random.default_rng(n)
    idx = np.concatenate([rng.integers(0, 1 << n, 2000, dtype=np.uint64), np.array([0, 1, 2, (1 << n) - 1, 1 << (n - 1)], dtype=np.uint64)])
    got = sv.amplitudes(idx)
    want = ghz_qft_closed_form(n, idx)
    scale = 2.0 ** (-n / 2)
    assert np.abs(got - want).max() < 1e-5 * scale, np.abs(got - want).max() / scale
    # distribution: p(y) = 2**-n (1 + cos(2 pi y / 2**n)); 64 bins over the top 6 outcome bits
    shots = 200_000
    outs = sv.sample(shots, list(range(n)), rng=np.random.default_rng(7))
    assert outs.max() < (1 << n)
    bins = (outs >> np.uint64(n - 6)).astype(np.int64)
    hist = np.bincount(bins, minlength=64)
    edges = np.arange(65) / 64.0
    cdf = edges + np.sin(2 * np.pi * edges) / (2 * np.pi)  # integral of 1 + cos(2 pi x)
    expect = shots * np.diff(cdf)
    keep = expect > 5
    chi2 = float(((hist[keep] - expect[keep]) ** 2 / expect[keep]).sum()) + float(hist[~keep].sum())
    assert chi2 < 120, chi2  # 63 degrees of freedom: P(chi2 > 120) ~ 1e-5


def test_multi_gpu_sharded_path():
    """Needs >= 2 GPUs on the box: launches tests/dist_gpu_check.py under torchrun."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "dist_gpu_check ok" in res.stdout


def test_backend_object_and_apply_matrix(sim):
    """The default-backend stand-in and the TensorFactor.apply_matrix stand-in."""
    from qrisp_b200.engine import StateVector

    qc = random_circuit(11, 4, seed=8)
    qc.measure([0, 3, 10])
    be = sim.B200Backend()
    got = be.run(qc, None)
    want = ref_sim.run(qc, None)
    assert set(got) == set(want) and max(abs(got[k] - want[k]) for k in got) < ATOL64
    # the reference's Backend contract (backend.py:217-330, qrisp_simulator_backend.py:33-110): run_async ->
    # a job that is done when it is handed out -> result().all_counts, one dict per circuit; a sequence of
    # circuits, per-circuit shots, options
    assert be.options == {"shots": None, "token": ""}
    job = be.run_async(qc)
    assert job.status() == "done" and job.cancel() is False
    assert job.result().num_circuits == 1 and job.result().get_counts() == got
    qc2 = random_circuit(9, 3, seed=2)
    qc2.measure(list(range(9)))
    both = be.run([qc, qc2], 500)
    assert isinstance(both, list) and len(both) == 2 and all(sum(d.values()) == 500 for d in both)
    per = be.run_async([qc, qc2], [100, 300]).result().all_counts
    assert [sum(d.values()) for d in per] == [100, 300]
    be.update_options(shots=64)
    assert sum(be.run(qc).values()) == 64
    with pytest.raises(AttributeError):
        be.update_options(nonsense=1)
    with pytest.raises(ValueError):
        be.run(qc, -5)
    # keyword compatibility with qrisp.simulator.run(qc, shots, token, iqs, insert_reset)
    assert sim.run(qc, None, "", None, True) == got and sim.run(qc, None, insert_reset=False) == got
    sv = StateVector(6)
    rng = np.random.default_rng(0)
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    u, _ = np.linalg.qr(a)
    sv.apply_matrix(u, (4, 1))
    want = ref_sim.apply_matrix(ref_sim.zero_state(6), u, (4, 1), 6, threshold=False)
    np.testing.assert_allclose(sv.statevector(), want, atol=ATOL64)


def test_qrisp_default_backend_on_the_b200(golden_dir):
    """BASELINE.json configs[0] end to end through Qrisp itself: with the vendored reference (oracle/_ref, see
    oracle/build_ref.py) importable, ``QuantumFloat(20)``, ``h``, ``QFT(inv=True)``, ``get_measurement()`` runs on
    the B200 backend installed as ``qrisp.default_backend.def_backend`` and gives what the reference's own
    simulator gave when the fixture was recorded."""
    from oracle import refstub

    qrisp = refstub.import_reference()
    if qrisp is None:
        pytest.skip("oracle/_ref is not built (python oracle/build_ref.py, in the build container)")
    import qrisp.default_backend as db
    from qrisp import QFT, QuantumFloat, h

    from qrisp_b200.backend import make_qrisp_backend

    g = np.load(os.path.join(golden_dir, "programs.npz"))
    old = db.def_backend
    db.def_backend = make_qrisp_backend()
    try:
        for size in (6, 20):
            qf = QuantumFloat(size)
            h(qf)
            QFT(qf, inv=True)
            mes = dict(qf.get_measurement(backend=db.def_backend))
            want = dict(zip((float(k) for k in g[f"qfloat_h_iqft{size}_decoded_keys"]), (float(v) for v in g[f"qfloat_h_iqft{size}_decoded_probs"])))
            assert set(float(k) for k in mes) == set(want), size
            assert max(abs(float(mes[k]) - want[float(k)]) for k in mes) < 1e-5
    finally:
        db.def_backend = old


def test_sparse_start_and_program_replay(sim):
    """The host prefix (qb2_sparse_prefix) in front of the dense sweeps: any cut gives the same state; a
    compiled program replays from a reset and refuses a map it was not compiled for."""
    from qrisp_b200 import _lib
    from qrisp_b200.engine import StateVector

    for qc in (ghz_qft_circuit(15), random_circuit(14, 5, seed=3, two_qubit="cx")):
        want = ref_sim.statevector(qc)
        instrs = [i for i in qc.data if i.matrix is not None]
        for support in (0, 1, 2, 16, 300, 1024):
            sv = StateVector(qc.num_qubits, prefix_support=support)
            sv.apply_circuit(instrs)
            np.testing.assert_allclose(sv.statevector(), want, atol=ATOL64, err_msg=f"support {support}")
            if support == 0:
                assert sv.stats.get("prefix_gates", 0) == 0
        sv = StateVector(qc.num_qubits)
        prog = sv.compile(instrs)
        for _ in range(2):
            sv.reset()
            prog.run()
            np.testing.assert_allclose(sv.statevector(), want, atol=ATOL64)
        with pytest.raises(_lib.Qb2Error):
            prog.run()  # the state is written and the map has moved: not what the program was compiled for
    # a circuit that the prefix finishes alone: nothing is launched until somebody looks at the state
    sv = StateVector(12)
    sv.apply_circuit([i for i in ghz_circuit(12).data if i.matrix is not None])
    assert sv.stats["passes"] == 0 and sv._pending is not None
    st = sv.statevector()
    assert abs(st[0] - 2**-0.5) < 1e-6 and abs(st[-1] - 2**-0.5) < 1e-6 and np.count_nonzero(st) == 2


def test_tma_kernel_matches_the_register_pipelined_kernel(sim):
    """Same circuits through pass_kernel_tma and pass_kernel: both against the oracle."""
    from qrisp_b200.engine import StateVector

    for seed, n in ((1, 13), (2, 16), (3, 19)):
        qc = random_circuit(n, 4, seed=seed, two_qubit="cx" if seed % 2 else "cz")
        qft_circuit(n, qc)
        want = ref_sim.statevector(qc)
        instrs = [i for i in qc.data if i.matrix is not None]
        used = {}
        for tma in (True, False):
            sv = StateVector(n, tma=tma, prefix_support=0)
            steps = sv.apply_circuit(instrs)
            used[tma] = sum(1 for s in steps if getattr(s, "tma", False))
            np.testing.assert_allclose(sv.statevector(), want, atol=ATOL64, err_msg=f"tma={tma} n={n}")
        assert used[True] > 0 and used[False] == 0


def test_shor_find_order_at_the_named_sizes(golden_dir, sim):
    """BASELINE.json configs[2]: the README's ``find_order`` for N = 35 (30 qubits, 16150 instructions), circuit as
    Qrisp compiled it (tests/golden/make_golden.py make_shor), through ``run()``: labels and probabilities equal
    the reference's.  (N = 77 -- 34 qubits, beyond the reference's own simulator -- is run by tools/shor_run.py
    and recorded under profiles/; minutes of device time do not belong into the test suite.)"""
    path = os.path.join(golden_dir, "shor.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/shor.npz not generated")
    import time

    g = np.load(path)
    cases = [str(c) for c in g["cases"] if str(c) + "_keys" in g.files]
    if not cases:
        pytest.skip("no case with a reference result in tests/golden/shor.npz")
    for case in cases:
        qc = Circuit.from_arrays(g, case + "_")
        t0 = time.perf_counter()
        got = sim.run(qc, None)
        secs = time.perf_counter() - t0
        ref = {str(k): float(p) for k, p in zip(g[case + "_keys"], g[case + "_probs"])}
        assert set(got) == set(ref), (case, len(got), len(ref))
        assert max(abs(got[k] - ref[k]) for k in ref) < 1e-5, case
        print(f"{case}: {qc.num_qubits} qubits, {len(qc.data)} instructions: B200 run() {secs:.2f} s, reference {float(g[case + '_ref_seconds']):.2f} s")


@pytest.mark.parametrize("switch", ["QB2_NO_SLOTS", "QB2_NO_CHAIN", "QB2_NO_HAD", "QB2_NO_RELABEL", "QB2_NO_BFLY"])
def test_planner_switches_do_not_change_the_state(sim, switch):
    """Every planner optimisation (phase slots, fused runs, unnormalised butterflies, relabelled
    stores, butterflies at all) can be switched off; the kernel's plain paths give the same state."""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); "
        "from qrisp_b200.circuit import ghz_qft_circuit; from qrisp_b200.simulator import statevector_sim; "
        "from oracle import ref_sim; qc = ghz_qft_circuit(20); "
        "e = np.abs(statevector_sim(qc) - ref_sim.statevector(qc)).max(); print('ERR', e); assert e < 1e-5"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **{switch: "1"})
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-1000:] + res.stderr[-2000:]


def test_fresh_state_is_zero_state_whoever_looks_first():
    """reset() is lazy (the first pass synthesises |0...0>); every other way of looking at the state
    has to find it written."""
    from qrisp_b200.engine import StateVector

    for n in (3, 11, 16):
        want = np.zeros(1 << n, dtype=np.complex64)
        want[0] = 1
        sv = StateVector(n)
        np.testing.assert_array_equal(sv.statevector(), want)
        sv = StateVector(n)
        raw = sv.state.cpu().numpy().view(np.complex64)  # the public buffer attribute writes the zeros too
        assert raw[0] == 1 and not raw[1:].any()
        sv = StateVector(n)
        assert abs(sv.norm2() - 1.0) < 1e-12
        sv = StateVector(n)
        p = sv.probabilities(list(range(min(n, 10))))
        assert p[0] == 1.0 and p[1:].sum() == 0.0
        sv = StateVector(n)
        assert set(sv.sample(64, rng=np.random.default_rng(0)).tolist()) == {0}
        # after gates, reset brings |0...0> back, through the first pass of the next circuit ...
        sv.apply_circuit([i for i in ghz_circuit(n).data if i.matrix is not None])
        sv.reset()  # |0...0> does not care where the qubits sit: the map may stay as it is
        sv.apply_circuit([i for i in ghz_circuit(n).data if i.matrix is not None])
        st = sv.statevector()
        assert abs(abs(st[0]) ** 2 - 0.5) < 1e-6 and abs(abs(st[-1]) ** 2 - 0.5) < 1e-6
        # ... or is written for real when the next thing is a read-out
        sv.reset()
        np.testing.assert_array_equal(sv.statevector(), want)


def test_random_gate_soup_on_the_device(sim):
    """Fuzz-style parity of the CUDA path: random mixes of gates (tests/soup.py: one- and two-qubit
    gates, Toffolis with mixed control states, controlled rotations, butterflies) on 9-16 qubits,
    several tile geometries, complex64 and complex128, against the oracle."""
    from soup import gate_soup

    rng = np.random.default_rng(77)
    geos = [{}, {"T": 9, "r": 3}, {"T": 10, "r": 4}, {"T": 12, "r": 5}, {"T": 13, "r": 4}, {"dtype": np.complex128}, {"dtype": np.complex128, "r": 4, "T": 11}]
    for case in range(140):
        n = int(rng.integers(9, 17))
        g = geos[case % len(geos)]
        if g.get("T", 0) > n:
            continue
        qc = gate_soup(n, int(rng.integers(5, 90)), rng, wide=case % 2 == 1)
        got = sim.statevector_sim(qc, **g)
        c128 = g.get("dtype") is np.complex128
        want = ref_sim.statevector(qc, dtype=np.complex128)
        np.testing.assert_allclose(got, want, atol=1e-11 if c128 else 3e-5, err_msg=f"case {case} n={n} {g}")


def test_random_measured_circuits_on_the_device(sim):
    """run(): random gate soups with terminal measurements of a random subset of qubits, some with a
    mid-circuit measurement and a reset in between (rewritten into CX onto fresh qubits,
    circuit_preprocessing.py:729-836); the outcome labels must be identical to the oracle's and
    the probabilities agree to 1e-5."""
    from soup import gate_soup

    rng = np.random.default_rng(4242)
    for case in range(30):
        n = int(rng.integers(3, 11))
        qc = gate_soup(n, int(rng.integers(3, 40)), rng, wide=case % 2 == 0) if n >= 3 else Circuit(n)
        if case % 3 == 0:
            q = int(rng.integers(0, n))
            qc.measure(q)
            if case % 6 == 0:
                qc.reset(q)
            tail = gate_soup(n, int(rng.integers(2, 15)), rng)
            qc.data.extend(tail.data)
        qubits = sorted(int(x) for x in rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False))
        qc.measure(qubits)
        got = sim.run(qc, None)
        want = ref_sim.run(qc, None)
        assert set(got) == set(want), (case, sorted(set(got) ^ set(want))[:5])
        assert max(abs(got[k] - want[k]) for k in got) < 1e-5, case


def test_dense_5q_blocks_on_the_tensor_cores(sim):
    """qb2_apply_block_tc (tcgen05.mma, TF32 x 3 split): Haar-random 5-qubit unitaries on low, high, mixed and
    unsorted target positions, with and without controls, against the oracle at the north-star tolerance;
    and against the SIMT path (qb2_apply_matrix_big) on the same circuit."""
    from qrisp_b200.engine import StateVector

    rng = np.random.default_rng(11)

    def haar(k):
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _r = np.linalg.qr(a)
        return q

    n = 14
    worst = 0.0
    for targets in ((0, 1, 2, 3, 4), (9, 10, 11, 12, 13), (13, 0, 7, 3, 10), (1, 5, 6, 8, 12), (4, 3, 2, 1, 0)):
        qc = random_circuit(n, 3, seed=sum(targets))
        qc.unitary(haar(5), targets)
        qc.h(2)
        free = [q for q in range(n) if q not in targets]
        qc.append("cu5", [free[0], free[3]] + list(targets), matrix=np.kron(np.diag([1, 0, 0, 0]), np.eye(32))
                  + np.kron(np.diag([0, 1, 1, 0]), np.eye(32)) + np.kron(np.diag([0, 0, 0, 1]), haar(5)))
        qc.cz(0, 13)
        want = ref_sim.statevector(qc)
        sv = StateVector(n)
        sv.apply_circuit([i for i in qc.data if i.matrix is not None])
        assert sv.stats.get("block_tc", 0) == 2, sv.stats
        got = sv.statevector()
        worst = max(worst, float(np.abs(got - want).max()))
        np.testing.assert_allclose(got, want, atol=ATOL64)
        simt = StateVector(n)
        simt.block_tc = False
        simt.apply_circuit([i for i in qc.data if i.matrix is not None])
        assert simt.stats.get("block_tc", 0) == 0
        np.testing.assert_allclose(simt.statevector(), want, atol=ATOL64)
    print(f"tensor-core 5-qubit blocks: worst |amp - oracle| = {worst:.2e}")


def test_specialised_pass_kernels_match_the_interpreter(sim):
    """qrisp_b200.jit: a pass compiled with its structure as compile-time constants (same kernel source) gives
    the interpreter's amplitudes -- and the oracle's -- on butterfly passes, random circuits and a sparse start;
    a second program with the same structure but other angles reuses the kernel."""
    from qrisp_b200 import jit
    from qrisp_b200.engine import StateVector

    if jit.find_nvcc() is None:
        pytest.skip("no nvcc on this box: the interpreter stays in charge")
    for qc, kw in ((ghz_qft_circuit(16), {}), (ghz_qft_circuit(17), {"prefix_support": 0}), (random_circuit(15, 4, seed=9), {})):
        instrs = [i for i in qc.data if i.matrix is not None]
        want = ref_sim.statevector(qc)
        plain = StateVector(qc.num_qubits, specialize=False, **kw)
        plain.apply_circuit(instrs)
        fast = StateVector(qc.num_qubits, specialize=True, **kw)
        fast.apply_circuit(instrs)
        assert fast.stats.get("jit_passes", 0) >= 1 and plain.stats.get("jit_passes", 0) == 0, fast.stats
        got = fast.statevector()
        np.testing.assert_allclose(got, plain.statevector(), atol=2e-7)
        np.testing.assert_allclose(got, want, atol=ATOL64)
    # complex128: the register-pipelined kernel's other geometry
    qc = random_circuit(13, 3, seed=5, two_qubit="cx")
    fast = StateVector(13, dtype=np.complex128, specialize=True)
    fast.apply_circuit([i for i in qc.data if i.matrix is not None])
    assert fast.stats.get("jit_passes", 0) >= 1
    np.testing.assert_allclose(fast.statevector(), ref_sim.statevector(qc, dtype=np.complex128), atol=ATOL128)
    # replayed program; same structure, other angles: one kernel
    a, b = random_circuit(15, 3, seed=1), random_circuit(15, 3, seed=1)
    for ins in b.data:
        if ins.name == "u3" and ins.matrix is not None:
            ins.matrix = ins.matrix @ np.diag([1, np.exp(0.3j)])
    keys = []
    for qc in (a, b):
        sv = StateVector(15, specialize=True, prefix_support=0)
        prog = sv.compile([i for i in qc.data if i.matrix is not None])
        assert prog.specialize() >= 1
        keys.append([jit.pass_key(jit.small_section(s.blob), s.tma) for s in prog.steps if s.kind == "pass"])
        sv.reset()
        prog.run()
        sv.reset()
        prog.run()
        np.testing.assert_allclose(sv.statevector(), ref_sim.statevector(qc), atol=ATOL64)
    assert keys[0] == keys[1]


# ---------------------------------------------------------------------------------------
# the callers of the path: programs from Qrisp's own repertoire (kept last in this file)
# ---------------------------------------------------------------------------------------


def test_qrisp_algorithms_match_reference_results(golden_dir, sim):
    """Grover, phase estimation, arithmetic and comparisons, logic synthesis, multi-controlled X by five
    methods, Dicke states, QAOA layers, modular multiplication, division, uncomputation, amplitude
    amplification -- compiled by Qrisp, captured at the backend boundary together with what the reference's
    ``run`` returned (tests/golden/make_golden.py: make_algorithms): same labels, probabilities within 1e-5."""
    g = load(golden_dir, "algorithms.npz")
    assert len(g["cases"]) >= 20
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        got = sim.run(qc, None)
        ref = {str(k): float(p) for k, p in zip(g[case + "_keys"], g[case + "_probs"])}
        assert set(got) == set(ref), case  # basis-state labels bit-exact
        for k in ref:
            assert abs(got[k] - ref[k]) < ATOL64, (case, k)
