"""`qb2_run_pass` checks a pass on the HOST before anything is launched (csrc/qb2_pass.cu: launch_pass,
validate_pass): the kernel reads the descriptors through the constant bank without bounds checks, so a
host / ABI mismatch must come back as QB2_EINVAL / QB2_ELIMIT, never as wrong amplitudes.  Without a device
the call ends with QB2_ENODEV once the checks are through, which is how this file tells "accepted" from
"rejected" -- no compute call is made."""

import ctypes
import struct

import numpy as np
import pytest
import torch

from qrisp_b200 import _lib
from qrisp_b200.circuit import Circuit, ghz_qft_circuit, random_circuit
from qrisp_b200.distributed import PlanOnlyShard
from qrisp_b200.planner import FEATURE_FULL, FEATURE_ZERO_INPUT

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="with a device an accepted pass would really be launched")

EINVAL, ENODEV, ELIMIT = -1, -3, -4
FAKE = ctypes.c_void_p(1 << 20)  # non-null, 16-byte aligned, never dereferenced on the host


def _passes(qc, n, world=1, **kw):
    steps, _entry, _pos = PlanOnlyShard(n, world, **kw).plan([i for i in qc.data if i.matrix is not None])
    return [s for s in steps if s.kind == "pass"]


def _call(small, n, dtype=0, blob_dev=FAKE, sparse=None, n_sparse=0):
    lib = _lib.load()
    hdr = ctypes.create_string_buffer(bytes(small), len(small))
    rc = lib.qb2_run_pass(FAKE, n, 0, hdr, blob_dev, dtype, None, sparse, n_sparse)
    return rc, lib.qb2_last_error().decode(errors="replace")


def _small(step):
    return bytearray(step.blob[: int.from_bytes(step.blob[8:12], "little")])


def test_validator_accepts_what_the_planner_builds(golden_dir):
    """Every pass of a spread of plans -- headline, random circuits with dense two-qubit gates, both dtypes,
    the register-pipelined geometries, sharded plans, and the Qrisp programs of algorithms.npz."""
    plans = [
        (ghz_qft_circuit(30), 30, 1, {}), (ghz_qft_circuit(33), 33, 8, {}), (ghz_qft_circuit(18), 18, 2, {"prefix_support": 2}),
        (random_circuit(20, 8, seed=1, two_qubit="cx"), 20, 1, {}), (random_circuit(17, 6, seed=2), 17, 4, {"tma": False}),
        (random_circuit(15, 4, seed=3), 15, 1, {"dtype": np.complex128}), (random_circuit(16, 4, seed=4), 16, 1, {"T": 11, "r": 4}),
    ]
    from soup import gate_soup  # every op kind the planner knows (dense 3/4-qubit matrices, general diagonals, ladders)

    for seed in range(12):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(10, 18))
        plans.append((gate_soup(n, 150, rng, wide=bool(seed % 2)), n, 1 + seed % 2, {"dtype": np.complex128} if seed % 3 == 0 else {}))
    g = np.load(golden_dir + "/algorithms.npz", allow_pickle=False)
    from qrisp_b200.preprocess import unitarize

    for case in g["cases"]:
        instrs, n, _ = unitarize(Circuit.from_arrays(g, str(case) + "_"))
        c = Circuit(n)
        c.data = instrs
        plans.append((c, max(n, 9), 1, {}))
    seen = 0
    for qc, n, world, kw in plans:
        sv = PlanOnlyShard(n, world, **kw)
        for s in _passes(qc, n, world, **kw):
            rc, msg = _call(_small(s), sv.n, dtype=sv.dtype)
            assert rc == ENODEV, (n, world, kw, rc, msg)
            seen += 1
    assert seen > 150


def test_validator_rejects_a_pass_that_does_not_fit():
    n = 16
    step = _passes(random_circuit(n, 6, seed=7, two_qubit="cx"), n)[0]
    good = _small(step)
    assert _call(good, n)[0] == ENODEV

    def with_u32(off, value):
        b = bytearray(good)
        struct.pack_into("<I", b, off, value)
        return b

    def with_u8(off, value):
        b = bytearray(good)
        b[off] = value
        return b

    small = len(good)
    phases_off, ops_off = struct.unpack_from("<II", good, 24)
    n_ops = struct.unpack_from("<I", good, 52)[0]
    assert n_ops >= 2
    cases = {
        "magic": (with_u32(0, 0xDEADBEEF), EINVAL),
        "state width": (with_u32(12, n + 1), EINVAL),
        "dtype": (with_u8(18, 1), EINVAL),
        "tile wider than the state": (with_u8(16, 40), ELIMIT),
        "no phases": (with_u32(20, 0), EINVAL),
        "phases beyond the small section": (with_u32(24, small - 16), EINVAL),
        "ops beyond the small section": (with_u32(28, small - 16), EINVAL),
        "misaligned coefficients": (with_u32(32, struct.unpack_from("<I", good, 32)[0] + 4), EINVAL),
        "more ops than there are": (with_u32(52, n_ops + 4000), EINVAL),
        "phase names ops outside the pass": (with_u32(phases_off + 4, n_ops + 1), EINVAL),
        "phase index arrays outside": (with_u32(phases_off + 8, 1 << 20), EINVAL),
        "unknown op kind": (with_u8(ops_off + 24, 77), EINVAL),
        "unknown dispatch code": (bytearray(good[: ops_off + 22]) + struct.pack("<H", 60000) + bytearray(good[ops_off + 24 :]), EINVAL),
        "register bit outside the phase": (with_u8(ops_off + 25, 9), EINVAL),
        "too many phase slots": (bytearray(good[:62]) + struct.pack("<H", 60000) + bytearray(good[64:]), ELIMIT),
    }
    for what, (blob, want) in cases.items():
        rc, msg = _call(blob, n)
        assert rc == want, (what, rc, msg)
        assert "qb2_run_pass" in msg, (what, msg)
    # pointer checks
    assert _call(good, n, blob_dev=ctypes.c_void_p((1 << 20) + 8))[0] == EINVAL  # blob not 16-byte aligned
    assert _call(good, n, blob_dev=None)[0] == EINVAL
    # a pass that starts from a sparse list takes at most QB2_MAX_SPARSE entries, and needs the list
    zero_in = with_u32(56, struct.unpack_from("<I", good, 56)[0] | FEATURE_ZERO_INPUT)
    assert _call(zero_in, n, sparse=FAKE, n_sparse=5000)[0] == ELIMIT
    assert _call(zero_in, n, sparse=None, n_sparse=8)[0] == EINVAL
    assert _call(zero_in, n, sparse=FAKE, n_sparse=8)[0] == ENODEV


def test_lean_pass_with_a_two_qubit_matrix_is_refused():
    """MAT2/3/4 and general DIAG arms are compiled out of the lean kernel variant: a pass that holds one must
    say QB2_FEATURE_FULL, or the op would silently be skipped."""
    n = 14
    qc = Circuit(n)
    rng = np.random.default_rng(0)
    for q in range(0, n - 1, 2):
        a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        qc.unitary(np.linalg.qr(a)[0], [q, q + 1])
    step = _passes(qc, n, prefix_support=0)[0]
    good = _small(step)
    features = struct.unpack_from("<I", good, 56)[0]
    assert features & FEATURE_FULL and _call(good, n)[0] == ENODEV
    lean = bytearray(good)
    struct.pack_into("<I", lean, 56, features & ~FEATURE_FULL)
    rc, msg = _call(lean, n)
    assert rc == EINVAL and "full kernel" in msg


def test_argument_checks_of_the_other_entry_points():
    """Bad arguments come back as QB2_EINVAL / QB2_ELIMIT with a message that names the function, before any
    device work (the reference raises Python exceptions at the same places: wrong qubit lists, shapes)."""
    lib = _lib.load()

    def ints(*v):
        return (ctypes.c_int * len(v))(*v)

    def refused(rc, who, code=EINVAL):
        msg = lib.qb2_last_error().decode(errors="replace")
        assert rc == code, (who, rc, msg)
        assert who in msg, (who, msg)

    other = ctypes.c_void_p(2 << 20)
    n = 20
    # layout
    refused(lib.qb2_permute_bits(FAKE, other, 4, ints(0, 1, 1, 3), 0, None), "qb2_permute_bits")  # not a permutation
    refused(lib.qb2_permute_bits(FAKE, FAKE, 4, ints(0, 1, 2, 3), 0, None), "qb2_permute_bits")  # in place
    peers = (ctypes.c_void_p * 2)(1 << 20, 2 << 20)
    refused(lib.qb2_swap_pull(other, peers, 3, 0, n, None, 0, None), "qb2_swap_pull")  # not a power of two
    refused(lib.qb2_swap_pull(other, peers, 2, 2, n, None, 0, None), "qb2_swap_pull")  # rank outside the world
    refused(lib.qb2_swap_pull(other, peers, 2, 0, 4, ints(1, 0, 2, 3), 0, None), "qb2_swap_pull")  # complex64 moves position 0
    refused(lib.qb2_swap_pull(other, peers, 2, 0, 4, ints(0, 2, 2, 3), 0, None), "qb2_swap_pull")
    peers[1] = None
    refused(lib.qb2_swap_pull(other, peers, 2, 0, 4, None, 0, None), "qb2_swap_pull")  # null peer
    # measurement
    refused(lib.qb2_probabilities(FAKE, n, 0, ints(3, 3), 2, other, 0, None), "qb2_probabilities")  # duplicate position
    refused(lib.qb2_probabilities(None, n, 0, ints(3), 1, other, 0, None), "qb2_probabilities")
    refused(lib.qb2_norm2(FAKE, 41, other, 0, None), "qb2_norm2")
    refused(lib.qb2_sample_scan(None, 16, None), "qb2_sample_scan")
    refused(lib.qb2_prune_probabilities(FAKE, 12, 2e-4, other, None, None, 8, None), "qb2_prune_probabilities")  # capacity without outputs
    # dense blocks
    refused(lib.qb2_apply_matrix_big(FAKE, other, n, 0, FAKE, ints(*range(11)), 11, 0, 0, 0, None), "qb2_apply_matrix_big", ELIMIT)
    refused(lib.qb2_apply_matrix_big(FAKE, other, n, 0, FAKE, ints(1, 1, 2), 3, 0, 0, 0, None), "qb2_apply_matrix_big")
    refused(lib.qb2_apply_matrix_big(FAKE, other, n, 0, FAKE, ints(1, 2, 3), 3, 1 << 2, 1 << 2, 0, None), "qb2_apply_matrix_big")  # control on a target
    refused(lib.qb2_apply_block_tc(FAKE, n, 0, other, ints(0, 1, 2, 4, 3), 0, 0, None), "qb2_apply_block_tc")  # not ascending
    refused(lib.qb2_apply_block_tc(FAKE, n, 0, other, ints(0, 1, 2, 3, 4), 1 << 3, 1 << 3, None), "qb2_apply_block_tc")  # control on a target
    refused(lib.qb2_apply_block_tc(FAKE, 8, 0, other, ints(0, 1, 2, 3, 4), 0, 0, None), "qb2_apply_block_tc")  # too narrow
    refused(lib.qb2_apply_block_tc(None, n, 0, other, ints(0, 1, 2, 3, 4), 0, 0, None), "qb2_apply_block_tc")
    refused(lib.qb2_apply_mux_ry(FAKE, n, 0, 3, ints(3), 1, other, 0, 0, 0, None), "qb2_apply_mux_ry")  # control = target
    refused(lib.qb2_apply_mux_ry(FAKE, n, 0, 3, ints(*range(4, 12)), 8, other, 0, 0, 0, None), "qb2_apply_mux_ry")  # too many controls
    refused(lib.qb2_pack_block_tc(None, None), "qb2_pack_block_tc")
    # specialised kernels
    refused(lib.qb2_jit_load(None, None), "qb2_jit_load")
    refused(lib.qb2_run_pass_jit(None, FAKE, n, 0, FAKE, FAKE, 0, None, None, 0), "qb2_run_pass_jit")
    # host prefix
    count, consumed = ctypes.c_uint32(5), ctypes.c_int32(0)
    refused(lib.qb2_sparse_prefix(0, None, None, None, None, 4, 1e-14, other, other, ctypes.byref(count), ctypes.byref(consumed)),
            "qb2_sparse_prefix")  # list longer than its capacity


def test_pack_block_tc_splits_into_two_tf32_halves():
    """Host half of the tensor-core block (qb2_pack_block_tc): the 32 x 32 complex matrix as the real 64 x 64
    operand with re / im interleaved per row and column ([[re, -im], [im, re]] per entry), each value as a TF32
    high part plus a TF32 low part, both in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
    matrices, 128 B between K chunks, 2 KB between row groups).  hi + lo gives the value back to 2**-21."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    a = rng.normal(size=(32, 32)) + 1j * rng.normal(size=(32, 32))
    u = np.ascontiguousarray(np.linalg.qr(a)[0].astype(np.complex128))
    out = np.zeros(_lib.QB2_BLOCK_TC_BYTES // 4, dtype=np.float32)
    assert lib.qb2_pack_block_tc(u.ctypes.data, out.ctypes.data) == 0
    assert not np.any(out.view(np.uint32) & np.uint32(0x1FFF)), "every stored word is a TF32 value (13 low mantissa bits clear)"
    hi, lo = out[: out.size // 2].astype(np.float64), out[out.size // 2 :].astype(np.float64)
    row, k = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    off = ((row >> 3) * 2048 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4) // 4
    assert sorted(off.ravel().tolist()) == list(range(64 * 64))  # the layout is a permutation of the tile
    entry = u[row >> 1, k >> 1]
    want = np.where(row & 1, np.where(k & 1, entry.real, entry.imag), np.where(k & 1, -entry.imag, entry.real))
    got = hi[off] + lo[off]
    assert np.abs(got - want).max() < 2.0**-21
    assert np.abs(hi[off] - want).max() < 2.0**-10 and np.abs(hi[off] - want).max() > 0  # hi alone is only TF32
