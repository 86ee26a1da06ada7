"""CPU tests of the host logic: gate normalisation, the pass planner and the packed pass
format, checked through the numpy pass emulator against the oracle and the golden fixtures
the reference produced."""

import os

import numpy as np
import pytest

from oracle import pass_emulator, ref_sim
from qrisp_b200 import preprocess, simulator
from qrisp_b200.circuit import Circuit, ghz_qft_circuit, random_circuit
from qrisp_b200.normalize import normalize_circuit, normalize_instruction

from helpers import canonical, emulate_statevector, marginal, plan_circuit


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------


def _ops_unitary(ops, qubits):
    """Dense unitary (big endian in ``qubits``) of a list of normalised ops."""
    k = len(qubits)
    dim = 1 << k
    u = np.eye(dim, dtype=np.complex128)
    bit_of = {q: k - 1 - i for i, q in enumerate(qubits)}  # index bit of each qubit
    perm = list(range(k))
    for op in ops:
        if op.kind == "relabel":
            g = np.zeros((dim, dim))
            a, b = bit_of[op.a], bit_of[op.b]
            for x in range(dim):
                xa, xb = (x >> a) & 1, (x >> b) & 1
                y = x & ~((1 << a) | (1 << b)) | (xb << a) | (xa << b)
                g[y, x] = 1
        elif op.kind == "diag":
            d = np.zeros(dim, dtype=np.complex128)
            for x in range(dim):
                idx = sum(((x >> bit_of[q]) & 1) << i for i, q in enumerate(op.qubits))
                d[x] = np.exp(2j * np.pi * op.turns[idx])
            g = np.diag(d)
        else:
            g = np.zeros((dim, dim), dtype=np.complex128)
            for x in range(dim):
                if all(((x >> bit_of[q]) & 1) == v for q, v in op.controls.items()):
                    col = sum(((x >> bit_of[q]) & 1) << i for i, q in enumerate(op.targets))
                    base = x
                    for q in op.targets:
                        base &= ~(1 << bit_of[q])
                    for row in range(1 << len(op.targets)):
                        y = base | sum(((row >> i) & 1) << bit_of[q] for i, q in enumerate(op.targets))
                        g[y, x] += op.matrix[row, col]
                else:
                    g[x, x] = 1
        u = g @ u
    return u


def test_normalize_every_golden_gate(golden_dir):
    g = load(golden_dir, "gates.npz")
    kinds = {}
    for name in g["names"]:
        name = str(name)
        u = g["op_" + name].astype(np.complex128)
        k = int(np.log2(u.shape[0]))
        qubits = tuple(int(q) for q in g["opq_" + name])
        ops = normalize_instruction(u, qubits)
        np.testing.assert_allclose(_ops_unitary(ops, list(range(k))), _embed(u, qubits, k), atol=1e-12, err_msg=name)
        kinds[name] = [o.kind for o in ops]
    assert kinds["swap"] == ["relabel"]
    for d in ("z", "s", "t", "p", "rz", "cz", "cp", "crz", "rzz", "mcp", "gphase"):
        assert kinds[d] == ["diag"], d
    assert kinds["cx"] == ["mat"] and kinds["mcx"] == ["mat"]
    # controls are recognised: cx / mcx have ONE dense target
    u = g["op_mcx"].astype(np.complex128)
    (op,) = normalize_instruction(u, (0, 1, 2))
    assert op.targets == (2,) and op.controls == {0: 1, 1: 1}


def _embed(u, qubits, k):
    """u (big endian in qubits) as a k-qubit matrix, big endian in range(k)."""
    assert sorted(qubits) == list(range(k))
    t = u.reshape(2 * k * [2])
    src = list(range(2 * k))
    dst = [qubits[i] for i in range(k)] + [k + qubits[i] for i in range(k)]
    return np.moveaxis(t, src, dst).reshape(1 << k, 1 << k)


def test_normalize_wide_and_uniformly_controlled():
    rng = np.random.default_rng(0)
    # 7-qubit multi-controlled phase: one CMat with six controls, not a 128-entry table
    d = np.ones(128, dtype=np.complex128)
    d[-1] = np.exp(0.3j)
    ops = normalize_instruction(np.diag(d), tuple(range(7)))
    assert [o.kind for o in ops] == ["mat"] and len(ops[0].controls) == 6
    np.testing.assert_allclose(_ops_unitary(ops, list(range(7))), np.diag(d), atol=1e-12)
    # uniformly controlled rotation: different block per control value
    blocks = []
    for _ in range(4):
        a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
        q, _r = np.linalg.qr(a)
        blocks.append(q)
    u = np.zeros((8, 8), dtype=np.complex128)
    for x in range(4):
        u[2 * x : 2 * x + 2, 2 * x : 2 * x + 2] = blocks[x]
    ops = normalize_instruction(u, (0, 1, 2))
    np.testing.assert_allclose(_ops_unitary(ops, [0, 1, 2]), u, atol=1e-12)
    # dense 3-qubit unitary stays one op
    a = rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8))
    q, _r = np.linalg.qr(a)
    ops = normalize_instruction(q, (2, 0, 1))
    assert len(ops) == 1 and len(ops[0].targets) == 3
    np.testing.assert_allclose(_ops_unitary(ops, [0, 1, 2]), _embed(q, (2, 0, 1), 3), atol=1e-12)


def test_normalize_circuit_fuses_one_qubit_runs():
    qc = Circuit(3)
    qc.h(0), qc.t(0), qc.rx(0.3, 0), qc.cx(0, 1), qc.h(1), qc.rxx(0.4, 1, 2), qc.h(2), qc.s(1)
    ops = normalize_circuit(qc.data)
    # h,t,rx -> one 1q op; h(1) is folded into rxx, h(2) and s(1) as well
    assert len(ops) == 3
    u = ref_sim.circuit_unitary(qc.data, [0, 1, 2])
    np.testing.assert_allclose(_ops_unitary(ops, [0, 1, 2]), u, atol=1e-12)


# ---------------------------------------------------------------------------------------
# planner + packed format through the emulator
# ---------------------------------------------------------------------------------------


def test_golden_statevectors_c64(golden_dir):
    g = load(golden_dir, "statevector.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        report = []
        sv, _ = emulate_statevector(qc, report=report)
        np.testing.assert_allclose(sv, g[case + "_sv"], atol=1e-5, err_msg=case)
        assert all(w == 1 for _, _, w in report), (case, report)


def test_golden_statevectors_c128(golden_dir):
    g = load(golden_dir, "statevector.npz")
    for case in ("rand5", "rand10", "ghzqft11"):
        qc = Circuit.from_arrays(g, case + "_")
        report = []
        sv, _ = emulate_statevector(qc, dtype=1, report=report)
        want = ref_sim.statevector(qc, dtype=np.complex128)
        np.testing.assert_allclose(sv, want, atol=1e-12, err_msg=case)
        assert all(w == 1 for _, _, w in report), (case, report)


def test_golden_fusion_circuits(golden_dir):
    g = load(golden_dir, "fusion.npz")
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        sv, _ = emulate_statevector(qc)
        np.testing.assert_allclose(sv, g[case + "_sv"], atol=1e-5, err_msg=case)


@pytest.mark.parametrize("T,r,L", [(9, 4, 4), (10, 3, 4), (11, 4, 5), (12, 4, 4), (13, 5, 4), (12, 5, 6), (13, 4, 4)])
def test_tile_geometries(T, r, L):
    qc = random_circuit(13, 4, seed=T * 10 + r, two_qubit="cx")
    qc.rxx(0.3, 0, 12)
    qc.mcx([1, 5, 9], 0)
    qc.swap(2, 11)
    qc.h(11)
    report = []
    sv, steps = emulate_statevector(qc, T=T, r=r, L=L, report=report)
    want = ref_sim.statevector(qc)
    np.testing.assert_allclose(sv, want, atol=2e-6)
    # exchanges are conflict free by construction; the tile buffers of the TMA kernel have the layout the
    # tensor map dictates: a distribution that holds tile bit 0 (or a swizzle pair) in registers pays 2-4 ways
    assert all(w == 1 for _, side, w in report if not side.startswith("tma")), report
    assert all(w <= 4 for _, side, w in report if side.startswith("tma")), report
    for s in steps:
        assert len(s.blob) % 256 == 0


def test_three_and_four_qubit_dense_ops_and_big_path():
    rng = np.random.default_rng(5)

    def haar(k):
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _r = np.linalg.qr(a)
        return q

    qc = Circuit(11)
    for q in range(11):
        qc.h(q)
    qc.unitary(haar(3), (7, 0, 4))
    qc.unitary(haar(4), (10, 2, 9, 5))
    qc.unitary(haar(3), (1, 2, 3))
    qc.unitary(haar(5), (0, 3, 6, 8, 10))  # wider than the register file: big path
    qc.cz(0, 1)
    sv, steps = emulate_statevector(qc)
    assert any(s.kind == "big" for s in steps)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc), atol=2e-6)
    sv5, steps5 = emulate_statevector(qc, T=13, r=5)
    np.testing.assert_allclose(sv5, ref_sim.statevector(qc), atol=2e-6)


def test_qft_needs_few_passes():
    """The point of the design: a 20-qubit GHZ+QFT (250 gates) is a handful of sweeps."""
    qc = ghz_qft_circuit(20)
    steps, pos, N = plan_circuit([i for i in qc.data if i.matrix is not None], 20)
    assert len(steps) <= 8, len(steps)
    # the final swaps of the QFT are relabels (the map is reversed, no data moved); passes that
    # end with low positions in registers store relabelled, which permutes the map further
    assert sorted(pos) == list(range(20))
    assert all(s.tma for s in steps)  # every tile of this circuit is one box of a tensor map
    import os
    # the register-pipelined kernel needs whole 128-byte lines on the lanes at the load and the store:
    # stored relabelled (the map moves), or with one more exchange
    os.environ["QB2_NO_TMA"] = "1"
    try:
        steps1, pos1, _ = plan_circuit([i for i in qc.data if i.matrix is not None], 20)
        os.environ["QB2_NO_RELABEL"] = "1"
        steps2, pos2, _ = plan_circuit([i for i in qc.data if i.matrix is not None], 20)
    finally:
        del os.environ["QB2_NO_TMA"]
        os.environ.pop("QB2_NO_RELABEL", None)
    assert not any(s.tma for s in steps1) and sorted(pos1) == list(range(20))
    assert pos2 == list(range(20))
    assert sum(s.n_phases for s in steps1) < sum(s.n_phases for s in steps2)  # fewer exchanges
    # the TMA kernel reads and writes its tile buffers under any distribution: no extra phases at all
    assert sum(s.n_phases for s in steps) <= sum(s.n_phases for s in steps1)
    assert pos == list(range(20))


def test_small_window_still_correct():
    qc = random_circuit(10, 6, seed=11)
    sv, _ = emulate_statevector(qc, window=7)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc), atol=2e-6)


def test_diag_tables_cover_nonlocal_positions():
    """Diagonal gates and controls on positions outside the tile (and on rank bits via
    hi_base) are evaluated from the physical index."""
    qc = Circuit(14)
    for q in range(14):
        qc.h(q)
    qc.cp(0.7, 0, 13)
    qc.cz(1, 2)
    qc.mcp(0.4, [0, 1, 12], 13)
    qc.cx(0, 13)  # control on the top position, target at the bottom
    qc.mcx([0, 1, 2], 12)
    sv, steps = emulate_statevector(qc)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc), atol=2e-6)
    # sharded: emulate 2 ranks of 13 local positions each, top qubit (= position 13) is the rank bit
    from qrisp_b200.normalize import normalize_circuit as nc
    from qrisp_b200.planner import PlanConfig, Planner

    instrs = [i for i in qc.data if i.matrix is not None][14:]  # after the H layer
    cfg = PlanConfig(13)
    pos = [13 - q for q in range(14)]
    steps, need = Planner(cfg, pos, n_total=14).plan(nc(instrs))
    assert need is None
    full = np.full(1 << 14, 2.0**-7, dtype=np.complex64)
    halves = []
    for rank in range(2):
        st = full[rank << 13 : (rank + 1) << 13].copy()
        for s in steps:
            st = pass_emulator.run(s.blob, st, hi_base=rank << 13)
        halves.append(st)
    got = canonical(np.concatenate(halves), pos, 14)
    np.testing.assert_allclose(got, ref_sim.statevector(qc), atol=2e-6)


# ---------------------------------------------------------------------------------------
# run(): circuit rewriting + outcome pruning (host logic), against the reference's results
# ---------------------------------------------------------------------------------------


def _run_emulated(qc):
    instrs, n, measurements = preprocess.unitarize(qc)
    c2 = Circuit(n)
    c2.data = instrs
    sv, _ = emulate_statevector(c2)
    measurements.sort(key=lambda x: -x[1])
    mes_qubits = [q for q, _ in measurements]
    p = marginal(sv, mes_qubits, n)
    outcomes, probs = simulator.prune_outcomes(p)
    probs = simulator.round_probabilities(probs)
    res = {}
    for o, pr in zip(outcomes, probs):
        if pr == 0:
            continue
        key = bin(int(o))[2:].zfill(len(mes_qubits))
        res[key] = res.get(key, 0.0) + float(pr)
    return res


@pytest.mark.parametrize("fixture", ["run.npz", "programs.npz", "cif.npz", "algorithms.npz"])
def test_run_matches_reference_results(golden_dir, fixture):
    g = load(golden_dir, fixture)
    for case in g["cases"]:
        case = str(case)
        qc = Circuit.from_arrays(g, case + "_")
        if qc.num_qubits > 16:
            continue  # the emulator is slow; the GPU test covers the wide cases
        got = _run_emulated(qc)
        ref = {str(k): float(p) for k, p in zip(g[case + "_keys"], g[case + "_probs"])}
        for k in set(got) | set(ref):
            assert abs(got.get(k, 0.0) - ref.get(k, 0.0)) < 1e-5, (case, k)


@pytest.mark.parametrize("fixture", ["run.npz", "cif.npz", "algorithms.npz"])
def test_unitarize_matches_oracle_rewriting(golden_dir, fixture):
    g = load(golden_dir, fixture)
    for case in g["cases"]:
        qc = Circuit.from_arrays(g, str(case) + "_")
        a = preprocess.unitarize(qc)
        b = ref_sim.rewrite_measurements(qc)
        assert a[1] == b[1] and a[2] == b[2]
        assert [(i.name, i.qubits) for i in a[0]] == [(i.name, i.qubits) for i in b[0]]


# ---------------------------------------------------------------------------------------
# the C-ABI library
# ---------------------------------------------------------------------------------------


def test_library_exports_every_declared_symbol():
    import re

    from qrisp_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "qrisp_b200.h")).read()
    declared = set(re.findall(r"\b(qb2_[a-z0-9_]+)\s*\(", header))
    declared -= {"qb2_run_pass_device"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.load()  # resolves every symbol, checks the ABI version
    assert lib.qb2_abi_version() == _lib.QB2_ABI_VERSION
    assert lib.qb2_sample_chunk_bits(30) == 12 and lib.qb2_sample_chunk_bits(8) == 8


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from qrisp_b200 import _lib
    from qrisp_b200.simulator import statevector_sim

    with pytest.raises(_lib.Qb2Error):
        statevector_sim(ghz_qft_circuit(3))


# ---------------------------------------------------------------------------------------
# the Qrisp-facing adapter (duck typed: Qrisp itself is not importable on the GPU box)
# ---------------------------------------------------------------------------------------


class _FakeOp:
    def __init__(self, name, matrix=None, params=()):
        self.name, self._m, self.params = name, matrix, params

    def get_unitary(self):
        return self._m


class _FakeInstr:
    def __init__(self, op, qubits, clbits=()):
        self.op, self.qubits, self.clbits = op, list(qubits), list(clbits)


class _FakeQC:
    """Looks like qrisp.QuantumCircuit as far as simulator.py:117-145 is concerned."""

    def __init__(self, n, ncl):
        self.qubits = [object() for _ in range(n)]
        self.clbits = [object() for _ in range(ncl)]
        self.data = []
        self.transpiled = False

    def transpile(self):
        res = _FakeQC(0, 0)
        res.qubits, res.clbits, res.data, res.transpiled = self.qubits, self.clbits, list(self.data), True
        return res


def test_from_qrisp_reads_the_fields_the_reference_reads():
    from qrisp_b200 import circuit as C

    qc = _FakeQC(3, 2)
    q, c = qc.qubits, qc.clbits
    qc.data.append(_FakeInstr(_FakeOp("qb_alloc"), [q[0]]))
    qc.data.append(_FakeInstr(_FakeOp("h", C.GATE_H.astype(np.complex64)), [q[2]]))
    qc.data.append(_FakeInstr(_FakeOp("cx", C._controlled(C.GATE_X)), [q[2], q[0]]))
    qc.data.append(_FakeInstr(_FakeOp("p", C.p_matrix(0.25), params=(0.25,)), [q[1]]))
    qc.data.append(_FakeInstr(_FakeOp("measure"), [q[0]], [c[1]]))
    qc.data.append(_FakeInstr(_FakeOp("measure"), [q[2]], [c[0]]))
    got = Circuit.from_qrisp(qc)
    assert got.num_qubits == 3 and got.num_clbits == 2
    assert [(i.name, i.qubits, i.clbits) for i in got.data] == [
        ("qb_alloc", (0,), ()), ("h", (2,), ()), ("cx", (2, 0), ()), ("p", (1,), ()), ("measure", (0,), (1,)),
        ("measure", (2,), (0,)),
    ]
    assert got.data[1].matrix.dtype == np.complex128 and got.data[3].params == (0.25,)
    # the oracle on the converted circuit: Bell pair on qubits (2, 0), clbit 0 <- qubit 2
    res = ref_sim.run(got, None)
    assert set(res) == {"00", "11"}
    with pytest.raises(NotImplementedError):
        bad = _FakeQC(1, 1)
        bad.data.append(_FakeInstr(_FakeOp("c_if_x", C.GATE_X), [bad.qubits[0]], [bad.clbits[0]]))
        Circuit.from_qrisp(bad)


def test_unnormalised_butterflies_settle_their_scale():
    """Hadamard-type butterflies leave their 1/sqrt(2) to the host (QB2_BFLY_HAD): whatever follows
    -- nothing, only relabels, a matrix too wide for a pass, more passes -- the state the steps
    leave behind is normalised."""
    import math

    from qrisp_b200.planner import BFLY_HAD, OP_BFLY

    def bfly_layer(qc, qubits):
        for q in qubits:
            qc.h(q)
            for k in range(q + 1, qc.num_qubits):
                qc.cp(math.pi / (1 << (k - q)), k, q)

    rng = np.random.default_rng(5)
    cases = {}
    qc = Circuit(12)
    bfly_layer(qc, range(12))
    cases["qft only"] = qc
    qc = Circuit(12)
    bfly_layer(qc, range(6))
    qc.swap(0, 11)
    qc.swap(3, 4)
    cases["relabels last"] = qc
    qc = Circuit(12)
    bfly_layer(qc, range(5))
    a = rng.normal(size=(32, 32)) + 1j * rng.normal(size=(32, 32))
    qc.unitary(np.linalg.qr(a)[0], (0, 3, 6, 8, 10))  # big path right after unnormalised butterflies
    bfly_layer(qc, range(5, 9))
    cases["big matrix in between"] = qc
    qc = Circuit(12)
    bfly_layer(qc, range(4))
    for q in range(12):
        qc.u3(0.3 + q, 0.2, 0.1, q)  # a general 2x2 absorbs the scale
    bfly_layer(qc, range(4, 12))
    cases["absorbed by a matrix"] = qc
    for name, qc in cases.items():
        for kw in ({}, {"T": 9, "r": 3}):
            sv, steps = emulate_statevector(qc, **kw)
            had = 0
            for s in steps:
                if s.kind == "pass":
                    had += sum(1 for op in pass_emulator.parse(s.blob).ops if op["kind"] == OP_BFLY and op["flags"] & BFLY_HAD)
            assert had > 0, name
            np.testing.assert_allclose(sv, ref_sim.statevector(qc), atol=3e-6, err_msg=name)


def test_planner_streams_steps_in_order():
    """plan(on_step=...) hands every step over as soon as it is built (the engine launches it while
    the next one is planned): same steps, same order, as the returned list."""
    from qrisp_b200.planner import PlanConfig, Planner

    qc = ghz_qft_circuit(18)
    ops = normalize_circuit([i for i in qc.data if i.matrix is not None])
    seen = []
    steps, need = Planner(PlanConfig(18, T=11, r=4), [17 - q for q in range(18)]).plan(ops, on_step=seen.append)
    assert need is None and len(steps) > 1
    assert [id(s) for s in seen] == [id(s) for s in steps]


def test_repeated_phase_gates_on_one_pair_add_up():
    """Regression: two controlled phases on the same pair of qubits (CZ twice, CP after CP) are one
    rung of a multiplier ladder with the SUM of the angles -- the ladder's mask can hold a partner
    position only once, so they must be merged before the ladder is formed."""
    qc = Circuit(11)
    qc.rxx(1.206, 9, 2)
    qc.cx(5, 8)
    qc.h(6)
    qc.cx(8, 0)
    qc.cz(9, 6)
    qc.cz(6, 9)
    sv, _ = emulate_statevector(qc, T=9, r=3)
    np.testing.assert_allclose(sv, ref_sim.statevector(qc, dtype=np.complex128), atol=3e-6)
    qc = Circuit(12)
    for q in range(12):
        qc.h(q)
    for a, b, th in ((1, 2, 0.785), (1, 2, 0.785), (4, 5, 3.644), (5, 4, 1.0), (7, 3, 0.1), (3, 7, 0.2), (7, 3, 0.3)):
        qc.cp(th, a, b)
    for q in (2, 5, 3):
        qc.h(q)
    for kw in ({}, {"T": 9, "r": 3}, {"T": 11, "r": 5}, {"dtype": 1}):
        sv, _ = emulate_statevector(qc, **kw)
        np.testing.assert_allclose(sv, ref_sim.statevector(qc, dtype=np.complex128), atol=3e-6)


def test_random_gate_soup_against_the_oracle():
    """A small fuzz run (the long one, ~17 000 circuits over seven geometries, found the bug above):
    random mixes of one- and two-qubit gates, Toffolis with mixed control states, controlled
    rotations, QFT-like butterflies (tests/soup.py) on 9-13 qubits."""
    from soup import gate_soup

    rng = np.random.default_rng(2024)
    geos = [{}, {"T": 9, "r": 3}, {"T": 10, "r": 4}, {"T": 11, "r": 5}, {"T": 9, "r": 4, "L": 4}, {"dtype": 1}]
    for case in range(72):
        n = int(rng.integers(9, 14))
        g = geos[case % len(geos)]
        if g.get("T", 0) > n:
            continue
        qc = gate_soup(n, int(rng.integers(5, 70)), rng, wide=case % 2 == 1)
        report = []
        sv, _ = emulate_statevector(qc, report=report, **g)
        tol = 3e-5 if g.get("dtype", 0) == 0 else 1e-11
        np.testing.assert_allclose(sv, ref_sim.statevector(qc, dtype=np.complex128), atol=tol, err_msg=f"case {case} {g}")
        assert all(w == 1 for _, side, w in report if not side.startswith("tma")), (case, report)


def test_reset_is_hoisted_over_z_commuting_gates():
    """measure q; cx(q, r); reset q: the look-ahead of the reference's insert_multiverse_measurements
    (circuit_preprocessing.py:755-789) takes the reset out of the stream at the measurement, so the CX that q only
    controls sees q AFTER the reset and never flips r.  The B200 rewriting reproduces that, and the
    oracle -- which is pinned to the reference's results on reset circuits (run.npz) -- agrees."""
    from oracle import ref_sim
    from qrisp_b200.circuit import Circuit
    from qrisp_b200.preprocess import unitarize

    qc = Circuit(2, 2)
    qc.x(0)
    qc.measure(0, 0)
    qc.cx(0, 1)
    qc.append("reset", [0])
    qc.measure(1, 1)
    instrs, n, meas = unitarize(qc)
    names = [i.name for i in instrs]
    assert names.index("cx") < len(names) - 1 and n == 3  # outcome copy and reset come first, the user's CX last
    want = ref_sim.run(qc, None)
    assert set(want) == {"01"}, want  # clbit 0 = 1 (measured before the reset), clbit 1 = 0: r was never flipped


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference: one JSON line with the keys of the bench contract (the CPU arm runs
    without a GPU; the sample register is shrunk here to keep the test short)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QB2_CPU_SAMPLE_QUBITS="16")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["unit"] == "amplitude-updates/s" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_cosine_sine_steps_rebuild_the_block():
    """Dense blocks on 6 and 7 qubits run as controlled 5-qubit blocks around multiplexed real rotations
    (engine._csd_steps, scipy's cosine-sine decomposition): the step list, applied with numpy under the kernels'
    index conventions (matrix index bit i <-> positions[i], controls as mask / value over the index), rebuilds
    the unitary."""
    from qrisp_b200.engine import DeviceOps

    class Host(DeviceOps):
        pass

    rng = np.random.default_rng(1)
    for k, positions in ((6, [2, 3, 5, 7, 8, 0]), (7, [0, 1, 2, 3, 4, 5, 6])):
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        u, _ = np.linalg.qr(a)
        n = max(positions) + 1
        steps = []
        Host()._csd_steps(u, positions, 0, 0, steps)
        assert sum(1 for s in steps if s[0] == "tc") == 4 ** (k - 5) and sum(1 for s in steps if s[0] == "mux") == (4 ** (k - 5) - 1) // 3
        dim = 1 << n
        idx = np.arange(dim)
        M = np.eye(dim, dtype=complex)
        for st in steps:
            on = ((idx ^ st[-1]) & st[-2]) == 0
            if st[0] == "tc":
                _, P, m, _cm, _cv = st
                row = np.zeros(dim, dtype=int)
                for t, p in enumerate(P):
                    row |= ((idx >> p) & 1) << t
                base = idx & ~sum(1 << p for p in P)
                acc = np.zeros_like(M)
                for c in range(32):
                    dep = sum(((c >> t) & 1) << p for t, p in enumerate(P))
                    acc += m[row, c][:, None] * M[base | dep, :]
                new = M.copy()
                new[on, :] = acc[on, :]
                M = new
            else:
                _, tg, ctrl, theta, _cm, _cv = st
                j = np.zeros(dim, dtype=int)
                for c, p in enumerate(ctrl):
                    j |= ((idx >> p) & 1) << c
                cth, sth = np.cos(theta)[j][:, None], np.sin(theta)[j][:, None]
                partner = M[idx ^ (1 << tg), :]
                new = np.where((((idx >> tg) & 1) == 0)[:, None], cth * M - sth * partner, sth * partner + cth * M)
                M = np.where(on[:, None], new, M)
        # the block acts on `positions` of an n-qubit register: compare with the embedded unitary
        row = np.zeros(dim, dtype=int)
        for t, p in enumerate(positions):
            row |= ((idx >> p) & 1) << t
        rest = idx & ~sum(1 << p for p in positions)
        want = np.where(rest[:, None] == rest[None, :], u[row[:, None], row[None, :]], 0)
        assert np.abs(M - want).max() < 1e-12


def _plan_sharded(nq, world, instrs):
    from qrisp_b200.distributed import PlanOnlyShard

    sv = PlanOnlyShard(nq, world)
    steps, entry, pos_in = sv.plan(instrs)
    return sv, dict(steps=steps, entry=entry, pos_in=pos_in)


def test_sharded_host_prefix_goes_world_times_further(monkeypatch):
    """Every rank takes QB2_MAX_SPARSE entries and the spread-out qubits start on the rank bits: GHZ + QFT
    stays at two passes when ranks widen the register (it needed three with the single-rank limit)."""
    from qrisp_b200 import prefix

    monkeypatch.delenv("QB2_PREFIX_SUPPORT", raising=False)
    for world, nq in ((1, 30), (2, 31), (4, 32), (8, 33)):
        qc = ghz_qft_circuit(nq)
        sv, got = _plan_sharded(nq, world, [i for i in qc.data if i.matrix is not None])
        kinds = [s.kind for s in got["steps"]]
        assert kinds == ["pass", "pass"], (world, kinds)
        idx, _ = got["entry"]
        assert 512 * world < len(idx) <= 1024 * world  # the walk really went further
        phys = prefix.to_physical(idx, got["pos_in"], nq)
        share = np.bincount((phys >> np.uint64(sv.n)).astype(np.int64), minlength=world)
        assert share.max() <= prefix.MAX_SUPPORT, (world, share)
    monkeypatch.setenv("QB2_PREFIX_SUPPORT", "1024")
    qc = ghz_qft_circuit(33)
    _, got = _plan_sharded(33, 8, [i for i in qc.data if i.matrix is not None])
    assert [s.kind for s in got["steps"]] == ["pass", "pass", "pass"]


def test_sharded_host_prefix_falls_back_when_a_rank_would_hold_too_much(monkeypatch):
    """The longer walk is checked against the layout: here the rank bit gets a qubit that no gate touches, so
    rank 0 would hold the whole list -- the walk is redone with the single-rank limit."""
    from qrisp_b200 import prefix

    monkeypatch.delenv("QB2_PREFIX_SUPPORT", raising=False)
    nq = 16
    qc = Circuit(nq)
    for q in range(12):
        qc.h(q)
    for q in range(11):
        qc.cx(q, q + 1)
    for q in range(12):
        qc.h(q)
    sv, got = _plan_sharded(nq, 2, list(qc.data))
    idx, _ = got["entry"]
    assert len(idx) <= prefix.MAX_SUPPORT
    assert sv.stats["prefix_gates"] == 10  # ten Hadamards, counted once
