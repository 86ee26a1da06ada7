"""Random gate soups for fuzz-style parity tests (shared by the CPU emulator test and the GPU test)."""
import math

import numpy as np

from qrisp_b200.circuit import Circuit, _controlled, ry_matrix, u3_matrix


def gate_soup(n, n_gates, rng, wide=False):
    """``n_gates`` random gates on ``n`` qubits: H, CP, CX, X, RZ, U3, SWAP, CZ, RY, RXX, QFT-like
    butterflies, and (``wide``) Toffoli / multi-controlled X with mixed control states, controlled
    RY / U3, CY, CRZ, MCP, Y, S, T, SX."""
    qc = Circuit(n)
    kinds = 20 if wide else 12
    for _ in range(n_gates):
        k = int(rng.integers(0, kinds))
        a, b, c = (int(x) for x in rng.choice(n, size=3, replace=False))
        th = float(rng.uniform(0, 2 * math.pi))
        if k == 0:
            qc.h(a)
        elif k == 1:
            qc.cp(th, a, b)
        elif k == 2:
            qc.cx(a, b)
        elif k == 3:
            qc.x(a)
        elif k == 4:
            qc.rz(th, a)
        elif k == 5:
            qc.u3(th, float(rng.uniform(0, 6.28)), float(rng.uniform(0, 6.28)), a)
        elif k == 6:
            qc.swap(a, b)
        elif k == 7:
            qc.cz(a, b)
        elif k == 8:
            qc.ry(th, a)
        elif k == 9:
            qc.cp(math.pi / 2 ** int(rng.integers(1, 6)), a, b)
        elif k == 10:
            qc.h(a)
            qc.cp(math.pi / 4, b, a)
            qc.cp(math.pi / 8, c, a)
        elif k == 11:
            qc.rxx(th, a, b)
        elif k == 12:
            qc.ccx(a, b, c)
        elif k == 13:
            qc.mcx([a, b], c, ctrl_state=int(rng.integers(0, 4)))
        elif k == 14:
            qc.unitary(_controlled(ry_matrix(th)), (a, b), name="cry")
        elif k == 15:
            qc.unitary(_controlled(u3_matrix(th, 0.4, 1.3)), (a, b), name="cu3")
        elif k == 16:
            qc.cy(a, b)
        elif k == 17:
            qc.crz(th, a, b)
        elif k == 18:
            qc.mcp(th, [a, b], c)
        else:
            [qc.y, qc.s, qc.t, qc.sx][int(rng.integers(0, 4))](a)
    return qc
