"""minidiff_b200: a B200-native engine with minidiff's API.

    import minidiff_b200 as md
    x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
    f = 2 * x * md.sin(x) - x**2
    f.backward()

Import order mirrors the reference's `minidiff/__init__.py:1-6` (backend first, then ops, then
tensor) but there is exactly one backend: hand-written sm_100a CUDA behind a C ABI.
"""
from . import backend  # noqa: F401  (raises if libminidiff_b200.so is missing: no CPU fallback)
from .tensor import *  # noqa: F401,F403,E402
from .tensor import Tensor, try_unwrap  # noqa: F401,E402
from .topology import OpNode  # noqa: F401,E402
from .ops.wrapping import *  # noqa: F401,F403,E402
from .ops.definitions import *  # noqa: F401,F403,E402
from .tensor import _install_operators as _install_operators  # noqa: E402
_install_operators(resolve=True)      # operator dunders call the op functions directly from here on
from .caching import reuse_graph  # noqa: F401,E402
from .graphs import CapturedGraph, capture_graph  # noqa: F401,E402
from .ops.fused_ops import make_ops as _make_fused_ops  # noqa: E402

import sys as _sys  # noqa: E402

# user-level fused device ops on the stateful-op protocol: extensions, not reference names
relu, linear_relu, linear = _make_fused_ops(_sys.modules[__name__])

__version__ = "0.1.0"
