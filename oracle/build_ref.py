"""Recipe for oracle/_ref: the UNMODIFIED reference simulator, runnable beside the B200 path.

    python oracle/build_ref.py            (also run by __graft_entry__.build())

Copies the reference's own Python package (``/root/reference/src/qrisp``) into ``oracle/_ref/qrisp``.
``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the GPU box like a built ``.so``.  Nothing is compiled and nothing is patched: the package is
imported through ``oracle.refstub``, which satisfies its module-scope imports of jax / qiskit /
matplotlib (absent in this image, and not on the simulator path) with inert stubs.

TEST INFRASTRUCTURE ONLY: ``bench.py --impl reference`` / ``cpu_baseline`` time it, tests compare against
it; nothing under ``qrisp_b200/`` imports it.
"""

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/qrisp"
DST = os.path.join(HERE, "_ref", "qrisp")


def build(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)  # the GPU box: use what travelled
    if os.path.isdir(DST) and not force:
        # refresh only when the reference tree is newer than the copy
        if os.path.getmtime(DST) >= max(os.path.getmtime(os.path.join(SRC, "simulator", f)) for f in os.listdir(os.path.join(SRC, "simulator")) if f.endswith(".py")):
            return True
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref/qrisp", "ready" if ok else "unavailable (no /root/reference here and no copy travelled)")
