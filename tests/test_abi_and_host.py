"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol declared in
include/minidiff_b200.h, the host-side view logic reproduces NumPy's shape/stride semantics, and the
product fails loudly -- never falls back -- when there is no device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import minidiff_b200.backend as B
from minidiff_b200.backend import _lib
from minidiff_b200.backend.device_array import DeviceArray, c_strides

HEADER = os.path.join(ROOT, "include", "minidiff_b200.h")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == syms, set(syms) ^ set(_lib.EXPORTS)
    assert lib.mdb_abi_version() == 1


def test_library_has_sm100a_code_and_no_cuda_link_dependency():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "libnccl" not in out      # static cudart, dlopen'd NCCL
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        lst = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "sm_100a" in lst


def test_backend_table_has_the_114_reference_names():
    names = B._F.EXPORTED
    assert len(names) == 114 and len(set(names)) == 114
    for n in names:
        assert hasattr(B, n), n
    assert B.tensor_class is DeviceArray
    assert B.float32 is np.float32 and B.bool is np.bool_
    assert B.sum.__name__ == "sum" and B.copy.__name__ == "copy"


@pytest.mark.skipif(_lib.device_available(), reason="only meaningful without a GPU")
def test_matmul_split_switch_is_host_only_state():
    """backend.set_matmul_split: the two operand splits of the tensor-core GEMM by name, anything else a ValueError
    (no device needed: the knob is host state of the library)."""
    B.set_matmul_split("fast")
    B.set_matmul_split("3xtf32")
    with pytest.raises(ValueError):
        B.set_matmul_split("tf32")
    for knob in (8, 9, 10):                      # SPLIT, CHUNK, RZ_GAIN back to their defaults
        _lib.check(_lib.lib.mdb_gemm_knob(knob, -1))


def test_no_cpu_fallback_without_a_device():
    with pytest.raises(RuntimeError, match="no CPU path"):
        B.ones((2, 2))
    import minidiff_b200 as md

    with pytest.raises(RuntimeError):
        md.Tensor([1.0, 2.0])


def test_missing_library_fails_at_import(tmp_path):
    code = ("import importlib.util, sys, os\n"
            "import minidiff_b200.backend._lib as L\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    # simulate by pointing the loader at a non-existent file through a patched copy of _lib.py
    src = open(os.path.join(ROOT, "minidiff_b200", "backend", "_lib.py")).read()
    src = src.replace('"lib", "libminidiff_b200.so"', '"lib", "definitely_missing.so"')
    p = tmp_path / "fake_lib.py"
    p.write_text(src.replace("_HERE = os.path.dirname(os.path.abspath(__file__))",
                             f"_HERE = {os.path.join(ROOT, 'minidiff_b200', 'backend')!r}"))
    r = subprocess.run([sys.executable, str(p)], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


# ---------------------------------------------------------------- host-side view semantics
def fake(shape, dtype=np.float32):
    """A DeviceArray over a fake pointer: views are metadata-only, so no device is needed."""
    a = np.zeros(shape, dtype)
    return DeviceArray(None, 1 << 20, a.shape, c_strides(a.shape), a.dtype), a


def eff(x):
    """strides that matter: those of axes with extent > 1 (NumPy's are arbitrary on unit axes)"""
    return tuple(st for e, st in zip(x.shape, x.strides) if e > 1)


VIEW_CASES = [
    ("transpose", (), {}), ("transpose", ((2, 0, 1),), {}), ("swapaxes", (0, 2), {}),
    ("expand_dims", (1,), {}), ("expand_dims", ((0, 4),), {}), ("squeeze", (), {}),
    ("reshape", ((6, 20),), {}), ("reshape", ((-1, 4),), {}), ("reshape", ((2, 3, 4, 5),), {}),
    ("ravel", (), {}), ("atleast_3d", (), {}), ("broadcast_to", ((2, 6, 4, 5),), {}),
]


@pytest.mark.parametrize("fn,args,kw", VIEW_CASES)
def test_view_metadata_matches_numpy(fn, args, kw):
    d, a = fake((6, 4, 5))
    got, want = getattr(B, fn)(d, *args, **kw), getattr(np, fn)(a, *args, **kw)
    assert got.shape == want.shape and eff(got) == eff(want), (got.strides, want.strides)


@pytest.mark.parametrize("key", [1, (2, 3), (slice(1, 5, 2), slice(None), -1), (Ellipsis, 0),
                                 (None, slice(None, None, -1)), (slice(5, 0, -2), None, 2),
                                 (slice(10, 20),), (1, Ellipsis, None)])
def test_basic_indexing_metadata_matches_numpy(key):
    d, a = fake((6, 4, 5))
    got, want = d[key], a[key]
    assert got.shape == want.shape and eff(got) == eff(want)
    off = (want.__array_interface__["data"][0] - a.__array_interface__["data"][0])
    if want.size:
        assert got.ptr - d.ptr == off


def test_reshape_of_noncontiguous_view_follows_numpy_nocopy_rule():
    d, a = fake((6, 4, 5))
    # mergeable: first two axes of a sliced-last-axis view
    v, w = d[:, :, ::2], a[:, :, ::2]
    got, want = B.reshape(v, (24, 3)), w.reshape(24, 3)
    assert eff(got) == eff(want) and got._st is v._st
    # transposes of contiguous arrays reshape without copy only when the rule allows it
    t, u = d.T, a.T
    assert eff(B.reshape(t, (5, 4, 6))) == eff(u.reshape(5, 4, 6))
    from minidiff_b200.backend.functions import _reshape_strides
    assert _reshape_strides(t.shape, t.estrides, [20, 6]) is None      # numpy would copy
    assert np.shares_memory(u.reshape(20, 6), a) is False


def test_flip_and_broadcast_flags():
    d, a = fake((3, 4))
    f, g = B.flip(d, axis=1), np.flip(a, axis=1)
    assert f.strides == g.strides and f.ptr - d.ptr == 3 * 4
    b = B.broadcast_to(d[0], (7, 4))
    assert b.strides == (0, 4) and not b.writeable
    with pytest.raises(ValueError):
        b += 1


def test_error_types_match_numpy():
    d, _ = fake((3, 4))
    with pytest.raises(IndexError):
        d[3]
    with pytest.raises(IndexError):
        d[0, 0, 0]
    with pytest.raises(ValueError):
        B.reshape(d, (5, 5))
    with pytest.raises(np.exceptions.AxisError):
        B.squeeze(d, axis=5)
    with pytest.raises(ValueError):
        B.broadcast_to(d, (4, 3))
    with pytest.raises(ValueError, match="could not be broadcast"):
        B._F.broadcast_shapes([(3, 4), (5, 4)])


def test_result_dtype_rules_follow_numpy():
    r = B._F.result_dtype
    f32, _ = fake((2,), np.float32)
    i64, _ = fake((2,), np.int64)
    b8, _ = fake((2,), np.bool_)
    f64, _ = fake((2,), np.float64)
    assert r(f32, 2) == np.float32 and r(f32, 2.5) == np.float32       # Python scalars are weak
    assert r(i64, 2.5) == np.float64 and r(i64, 2) == np.int64
    assert r(f32, i64) == np.float64 and r(f32, b8) == np.float32
    assert r(f32, f64) == np.float64 and r(b8, 1) == np.int64
    assert r(f32, np.float64(2)) == np.float64                          # NumPy scalars are strong


def test_engine_graph_bookkeeping_without_device():
    """Tensor / OpNode bookkeeping is host logic: exercise it on fake storage."""
    import minidiff_b200 as md

    d, _ = fake((2, 3))
    t = md.Tensor(d, allow_grad=True)
    assert t.shape == (2, 3) and t.is_leaf and not t.graphed and t.dtype == np.float32
    v = md.transpose(t)                                   # view op: records a node, no kernel
    assert v.shape == (3, 2) and not v.is_leaf and t.graph_refs == 1 and v.allow_grad
    assert v.op_node.op_name == "transpose" and v.op_node.toposort() == [t]
    with md.no_grad():
        w = md.reshape(t, (3, 2))
    assert w.is_leaf and not w.allow_grad
    with pytest.raises(ValueError):
        md.transpose(d)                                   # ops need Tensors (wrapping.py:28-44)
    with pytest.raises(ValueError):
        v.allow_grad = False
