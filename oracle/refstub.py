"""Import the reference simulator (``oracle/_ref/qrisp``, see ``oracle/build_ref.py``) in an image
without jax / qiskit / matplotlib.  TEST INFRASTRUCTURE ONLY.

The reference imports those packages at module scope; none of them is on the path from a
``QuantumCircuit`` to ``qrisp.simulator.run`` / ``statevector_sim``.  A meta-path finder hands out inert
stub modules for them (``jax.numpy`` is mapped to numpy, which is all ``QuantumFloat.decoder`` needs).
The reference code itself is untouched.

    from oracle import refstub
    qrisp = refstub.import_reference()        # None when oracle/_ref is absent
"""

import importlib.abc
import importlib.machinery
import os
import sys
import types

STUB_ROOTS = ("jax", "jaxlib", "qiskit", "matplotlib", "stim", "xdsl", "catalyst", "pennylane")
REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:  # decorator use: hand the function back
            return a[0]
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, o):
        return self

    def __ror__(self, o):
        return self


class _DummyMeta(type):
    def __call__(cls, *a, **k):
        if "_is_stub" not in cls.__dict__:
            return type.__call__(cls, *a, **k)
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()

    def __or__(cls, o):
        return cls

    def __ror__(cls, o):
        return cls

    def __getitem__(cls, k):
        return cls

    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()


def _stub_getattr(self, n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Dummy()


class StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        # classes, so that isinstance(x, jax.core.Tracer) and subclassing work
        cls = _DummyMeta(name, (), {"_is_stub": True, "__init__": lambda self, *a, **k: None,
                                    "__call__": lambda self, *a, **k: _Dummy(), "__getattr__": _stub_getattr})
        setattr(self, name, cls)
        return cls


class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_qrisp = None


def import_reference(ref_dir=None):
    """The reference's ``qrisp`` package, or None when no copy is available."""
    global _qrisp
    if _qrisp is not None:
        return _qrisp
    ref_dir = REF_DIR if ref_dir is None else ref_dir
    if not os.path.isdir(os.path.join(ref_dir, "qrisp")):
        return None
    if not any(isinstance(f, StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, StubFinder())
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import numpy

    import jax  # the stub
    import jax._src.core
    import jax.core

    sys.modules["jax.numpy"] = numpy
    jax.numpy = numpy
    ns = types.SimpleNamespace(trace=None)
    jax._src.core.trace_ctx = ns
    jax.core.trace_ctx = ns
    import qrisp

    _qrisp = qrisp
    return qrisp
