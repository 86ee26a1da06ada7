"""Import the reference with permissive stubs for absent deps (jax, qiskit, matplotlib)."""
import sys, types, importlib.abc, importlib.machinery

STUB_ROOTS = ("jax", "jaxlib", "qiskit", "matplotlib", "stim", "xdsl", "catalyst", "pennylane")

class _Dummy:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k):
        # decorator-friendly: return the function if used as decorator
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()
    def __getitem__(self, k): return _Dummy()
    def __iter__(self): return iter(())
    def __mro_entries__(self, bases): return (object,)
    def __or__(self, o): return self
    def __ror__(self, o): return self

class _DummyMeta(type):
    def __call__(cls, *a, **k):
        if "_is_stub" not in cls.__dict__:
            return type.__call__(cls, *a, **k)
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()
    def __or__(cls, o): return cls
    def __ror__(cls, o): return cls
    def __getitem__(cls, k): return cls
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

class StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        # classes so that isinstance(x, jax.core.Tracer) works and subclassing works
        cls = _DummyMeta(name, (), {"_is_stub": True, "__init__": lambda self,*a,**k: None, "__call__": lambda self,*a,**k: _Dummy(),
                                    "__getattr__": lambda self, n: (_ for _ in ()).throw(AttributeError(n)) if n.startswith("__") else _Dummy()})
        setattr(self, name, cls)
        return cls

class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None
    def create_module(self, spec):
        m = StubModule(spec.name); m.__path__ = []; return m
    def exec_module(self, module): pass

def install(ref_src="/root/reference/src"):
    sys.meta_path.insert(0, StubFinder())
    sys.path.insert(0, ref_src)

def patch():
    import types as _t
    import numpy
    import jax, jax.core, jax._src.core
    sys.modules["jax.numpy"] = numpy
    jax.numpy = numpy
    ns = _t.SimpleNamespace(trace=None)
    jax._src.core.trace_ctx = ns
    jax.core.trace_ctx = ns
