"""Specialised pass kernels (qrisp_b200/jit.py) without a GPU: the generated decode schedule is the interpreter's
walk over the op records, the cache key depends on structure and not on angles, and the generated translation
units compile for sm_100a (nvcc cross-compiles here) for both kernel bodies."""

import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import plan_circuit
from oracle import pass_emulator
from qrisp_b200 import jit
from qrisp_b200.circuit import ghz_qft_circuit, random_circuit


def _passes(qc, n, dtype=0):
    steps, _pos, _N = plan_circuit([i for i in qc.data if i.matrix is not None], n, dtype=dtype)
    return [s for s in steps if s.kind == "pass"]


def _walk(P):
    """Decode points per phase, from the emulator's independent reading of the format."""
    shapes = [(bs, k) for bs in range(1, 5) for k in range(2, bs + 2)]
    cxl = [(bs, k) for bs in range(2, 5) for k in range(2, bs + 1)]
    out = []
    for ph in P.phases:
        o, pts = ph[0], []
        while o < ph[0] + ph[1]:
            pts.append(o)
            code = P.ops[o]["code"]
            if 131 <= code < 141:
                o += shapes[code - 131][1]
            elif code == 141:
                o += bin(P.ops[o]["tab"]).count("1")
            elif 142 <= code < 148:
                o += cxl[code - 142][1]
            else:
                o += 1
        out.append(pts)
    return out


def test_decode_schedule_is_the_interpreters_walk():
    for qc, n in ((ghz_qft_circuit(16), 16), (random_circuit(15, 6, seed=4, two_qubit="cx"), 15), (ghz_qft_circuit(20), 20)):
        for s in _passes(qc, n):
            small = jit.small_section(s.blob)
            assert jit.decode_schedule(small) == _walk(pass_emulator.parse(s.blob))
            src = jit.kernel_source(small, s.tma)
            assert "jit_op(" in src and ("pass_tma_body" if s.tma else "pass_body") in src
            files = jit.schedule_files(small)
            assert "jit_phases.inc" in files and sum(f.count("qb2_pass_ops.inc") for f in files.values()) == sum(
                len(p) for p in jit.decode_schedule(small))


def test_cache_key_ignores_angles_but_not_structure():
    a = random_circuit(14, 3, seed=1)
    b = random_circuit(14, 3, seed=1)
    for ins in b.data:
        if ins.name == "u3" and ins.matrix is not None:
            ins.matrix = ins.matrix @ np.diag([1, np.exp(0.3j)])
    ka = [jit.pass_key(jit.small_section(s.blob), s.tma) for s in _passes(a, 14)]
    kb = [jit.pass_key(jit.small_section(s.blob), s.tma) for s in _passes(b, 14)]
    assert ka == kb
    kc = [jit.pass_key(jit.small_section(s.blob), s.tma) for s in _passes(random_circuit(14, 3, seed=2), 14)]
    assert ka != kc
    assert ka[0] != jit.pass_key(jit.small_section(_passes(a, 14)[0].blob), not _passes(a, 14)[0].tma)


@pytest.mark.skipif(jit.find_nvcc() is None, reason="no nvcc")
def test_specialised_translation_units_compile_for_sm_100a(tmp_path, monkeypatch):
    monkeypatch.setenv("QB2_JIT_CACHE", str(tmp_path))
    seen = set()
    # a complex64 TMA pass and a complex128 pass of the register-pipelined kernel (another geometry)
    for qc, n, dtype in ((ghz_qft_circuit(16), 16, 0), (random_circuit(15, 2, seed=4), 15, 1)):
        for s in _passes(qc, n, dtype):
            if s.tma in seen:
                continue
            seen.add(s.tma)
            path = jit.build_cubin(jit.small_section(s.blob), s.tma)
            assert path and os.path.getsize(path) > 10000
            cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
            if os.path.exists(cuobjdump):
                res = subprocess.run([cuobjdump, "-res-usage", path], capture_output=True, text=True).stdout
                assert "qb2_pass_jit" in res and "STACK:0" in res.split("qb2_pass_jit")[1], res  # registers stay registers
    assert seen == {True, False}


@pytest.mark.skipif(jit.find_nvcc() is None, reason="no nvcc")
def test_prebuild_fills_the_cache_without_a_device(tmp_path, monkeypatch):
    """What ``__graft_entry__.build()`` does for the bench workload, on a small circuit: plan as the ranks would
    (no device, no process group), cross-compile one kernel per distinct pass structure."""
    monkeypatch.setenv("QB2_JIT_CACHE", str(tmp_path))
    qc = ghz_qft_circuit(22)
    instrs = [i for i in qc.data if i.matrix is not None]
    passes, cubins = jit.prebuild(instrs, 22, world=2)
    assert passes >= 1 and cubins == passes
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")]) == passes
    # a second call finds everything in the cache; another world size with the same pass structures adds nothing
    assert jit.prebuild(instrs, 22, world=2) == (passes, cubins)
