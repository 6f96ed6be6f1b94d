"""Logging backend plugin for the REAL reference (test infrastructure; only used by
oracle/make_golden.py in the build container).  Selected through the reference's own loader
(`--backend oracle._trace_backend`, minidiff/backend/__init__.py:13-59); wraps every callable of the
reference's NumPy backend so the sequence of backend calls per config can be recorded."""
import minidiff.backend as backend
import minidiff.backend.numpy as _nb

CALLS = []
_QUIET = {"tensor_shape", "tensor_size", "tensor_ndim", "tensor_dtype", "dtype"}


def _logged(name, fn):
    def call(*a, **k):
        if name not in _QUIET:
            CALLS.append(name)
        return fn(*a, **k)

    call.__name__ = name
    return staticmethod(call)


_ns = {}
for _k, _v in vars(_nb.numpy_backend).items():
    if _k.startswith("_"):
        continue
    _f = getattr(_nb.numpy_backend, _k)
    if callable(_f) and not isinstance(_f, type):
        _ns[_k] = _logged(_k, _f)
    else:
        _ns[_k] = _f

tracing_backend = type("tracing_backend", (backend.Backend,), _ns)
del _nb  # keep exactly one Backend subclass reachable from this module's dict
