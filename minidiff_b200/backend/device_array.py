"""DeviceArray: the `tensor_class` of the B200 backend.

A strided view (device pointer + shape + element strides + dtype) over a refcounted allocation from
the library's caching allocator.  It provides what the reference expects of raw backend arrays
*outside* the function table (SURVEY 8b "Raw-array protocol"): `.astype`, the `+= -= *= /= //= **=
%= @=` family (reference tensor.py:269-362 mutates `_data` directly), `data[key] = value`
(tensor.py:376-379), `.size`, `.item()`, `__array__` (and *no* `__array_interface__`, so NumPy
falls through to `__array__`, tensor.py:424-433).

View operations (transpose / reshape / broadcast_to / slicing / flip ...) are metadata-only here,
with NumPy's exact shape & stride semantics; nothing is launched for them.
"""
from __future__ import annotations

import ctypes as C
import math
import operator

import numpy as np

from . import _lib
from ._lib import MdbArray, check, lib

_DT_CODE = {
    np.dtype(np.bool_): _lib.BOOL, np.dtype(np.uint8): _lib.U8, np.dtype(np.int8): _lib.I8,
    np.dtype(np.int16): _lib.I16, np.dtype(np.int32): _lib.I32, np.dtype(np.int64): _lib.I64,
    np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.uint16): _lib.U16,
    np.dtype(np.uint32): _lib.U32, np.dtype(np.uint64): _lib.U64, np.dtype(np.float16): _lib.F16,
}
F32 = np.dtype(np.float32)
F64 = np.dtype(np.float64)
I64 = np.dtype(np.int64)
BOOL = np.dtype(np.bool_)


def dtype_code(dt) -> int:
    try:
        return _DT_CODE[dt]
    except KeyError:
        raise TypeError(f"dtype {dt} is not supported by the B200 backend") from None


class _Storage:
    """One allocation from the caching allocator; freed (returned to the cache, stream-ordered)
    when the last view dies."""

    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, nbytes: int, ptr=None):
        if ptr is None:
            _lib.ensure_device()
            p = C.c_void_p()
            check(lib.mdb_alloc(max(int(nbytes), 1), C.byref(p)))
            ptr = p.value
        self.ptr = ptr          # adopted: allocated by the library inside mdb_elementwise_new
        self.nbytes = nbytes

    def __del__(self, _free=lib.mdb_free):
        try:
            _free(self.ptr)
        except Exception:  # interpreter shutdown
            pass


def c_strides(shape):
    st, acc = [0] * len(shape), 1
    for i in range(len(shape) - 1, -1, -1):
        st[i] = acc
        acc *= shape[i]
    return tuple(st)


class DeviceArray:
    __slots__ = ("_st", "ptr", "shape", "estrides", "dtype", "writeable", "_desc", "size",
                 "__weakref__")

    def __init__(self, storage, ptr, shape, estrides, dtype, writeable=True):
        self._st = storage
        self.ptr = ptr
        self.shape = shape
        self.estrides = estrides
        self.dtype = dtype
        self.writeable = writeable
        self._desc = None
        self.size = math.prod(shape)

    # ------------------------------------------------------------------ construction
    @classmethod
    def empty(cls, shape, dtype=F32) -> "DeviceArray":
        shape = tuple(int(s) for s in shape)
        if len(shape) > _lib.MAX_DIMS:
            raise ValueError(f"at most {_lib.MAX_DIMS} dimensions are supported")
        dtype = np.dtype(dtype)
        dtype_code(dtype)
        st = _Storage(math.prod(shape) * dtype.itemsize)
        return cls(st, st.ptr, shape, c_strides(shape), dtype)

    @classmethod
    def _adopt(cls, desc, shape, estrides, dtype, size) -> "DeviceArray":
        """Wrap an output the library allocated itself (mdb_elementwise_new filled desc.ptr); the
        descriptor becomes the array's cached one.  Hot path: no argument normalisation."""
        ptr = desc.ptr
        self = cls.__new__(cls)
        self._st = _Storage(size * dtype.itemsize, ptr)
        self.ptr = ptr
        self.shape = shape
        self.estrides = estrides
        self.dtype = dtype
        self.writeable = True
        self._desc = desc
        self.size = size
        return self

    @classmethod
    def from_numpy(cls, a) -> "DeviceArray":
        a = np.asarray(a)
        if a.dtype == object:
            raise TypeError("object arrays cannot live on the device")
        host = np.ascontiguousarray(a).reshape(a.shape)   # ascontiguousarray promotes 0-d to 1-d
        out = cls.empty(host.shape, host.dtype)
        if host.size:
            check(lib.mdb_h2d(out.ptr, host.ctypes.data, host.nbytes))
        return out

    def view(self, shape, estrides, ptr=None, writeable=None) -> "DeviceArray":
        return DeviceArray(self._st, self.ptr if ptr is None else ptr, tuple(shape),
                           tuple(estrides), self.dtype,
                           self.writeable if writeable is None else writeable)

    # ------------------------------------------------------------------ descriptors
    @property
    def d(self) -> MdbArray:
        d = self._desc
        if d is None:
            d = MdbArray()
            d.ptr = self.ptr
            d.dtype = _DT_CODE[self.dtype]
            n = d.ndim = len(self.shape)
            for i in range(n):
                d.shape[i] = self.shape[i]
                d.strides[i] = self.estrides[i]
            self._desc = d
        return d

    # ------------------------------------------------------------------ numpy-like metadata
    ndim = property(lambda s: len(s.shape))
    itemsize = property(lambda s: s.dtype.itemsize)
    nbytes = property(lambda s: s.size * s.dtype.itemsize)
    strides = property(lambda s: tuple(e * s.dtype.itemsize for e in s.estrides))
    base = property(lambda s: s._st)

    @property
    def T(self):
        return self.view(self.shape[::-1], self.estrides[::-1])

    @property
    def __array_interface__(self):
        raise AttributeError("device memory has no host array interface; use __array__/as_numpy")

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, not self.writeable),
                "version": 3, "strides": None if self.is_c_contiguous() else self.strides}

    def is_c_contiguous(self) -> bool:
        acc = 1
        for s, e in zip(reversed(self.shape), reversed(self.estrides)):
            if s != 1 and e != acc:
                return False
            acc *= s
        return True

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    # ------------------------------------------------------------------ host transfer (syncs)
    def numpy(self) -> np.ndarray:
        src = self if self.is_c_contiguous() else F.copy_(self)
        out = np.empty(self.shape, self.dtype)
        if out.size:
            check(lib.mdb_d2h(out.ctypes.data, src.ptr, out.nbytes))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None or np.dtype(dtype) == a.dtype else a.astype(dtype)

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.numpy().reshape(()).item()

    def tolist(self):
        return self.numpy().tolist()

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __index__(self):
        if not np.issubdtype(self.dtype, np.integer):
            raise TypeError("only integer scalar arrays can be converted to a scalar index")
        return int(self.item())

    def __bool__(self):
        if self.size != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous. "
                             "Use a.any() or a.all()")
        return bool(self.item())

    def __repr__(self):
        return "b200_" + repr(self.numpy())

    __str__ = lambda self: str(self.numpy())

    def __iter__(self):
        if not self.shape:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    # ------------------------------------------------------------------ methods (forward to F)
    def astype(self, dtype, copy=True):
        return F.astype(self, dtype, copy=copy)

    def copy(self):
        return F.copy_(self)

    def reshape(self, *shape, order="C"):
        if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
            shape = shape[0]
        return F.reshape(self, shape, order=order)

    def transpose(self, *axes):
        if len(axes) == 1 and not isinstance(axes[0], (int, np.integer)):
            axes = axes[0]
        return F.transpose(self, axes if axes else None)

    def ravel(self, order="C"):
        return F.ravel(self, order=order)

    def flatten(self, order="C"):
        return F.flatten(self, order=order)

    def squeeze(self, axis=None):
        return F.squeeze(self, axis=axis)

    def sum(self, axis=None, keepdims=False):
        return F.sum_(self, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, keepdims=False):
        return F.mean(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, keepdims=False):
        return F.max_(self, axis=axis, keepdims=keepdims)

    def min(self, axis=None, keepdims=False):
        return F.min_(self, axis=axis, keepdims=keepdims)

    def any(self, axis=None, keepdims=False):
        return F.any_(self, axis=axis, keepdims=keepdims)

    def all(self, axis=None, keepdims=False):
        return F.all_(self, axis=axis, keepdims=keepdims)

    def fill(self, value):
        self._check_writeable()
        check(lib.mdb_fill(C.byref(self.d), float(value)))

    # ------------------------------------------------------------------ operators
    def __add__(s, o): return F.add(s, o)
    def __radd__(s, o): return F.add(o, s)
    def __sub__(s, o): return F.subtract(s, o)
    def __rsub__(s, o): return F.subtract(o, s)
    def __mul__(s, o): return F.multiply(s, o)
    def __rmul__(s, o): return F.multiply(o, s)
    def __truediv__(s, o): return F.true_divide(s, o)
    def __rtruediv__(s, o): return F.true_divide(o, s)
    def __floordiv__(s, o): return F.floor_divide(s, o)
    def __rfloordiv__(s, o): return F.floor_divide(o, s)
    def __mod__(s, o): return F.mod(s, o)
    def __rmod__(s, o): return F.mod(o, s)
    def __pow__(s, o): return F.power(s, o)
    def __rpow__(s, o): return F.power(o, s)
    def __matmul__(s, o): return F.matmul(s, o)
    def __rmatmul__(s, o): return F.matmul(o, s)
    def __neg__(s): return F.negative(s)
    def __pos__(s): return F.copy_(s)
    def __abs__(s): return F.absolute(s)
    def __invert__(s): return F.invert(s)
    def __eq__(s, o): return F.equal(s, o)
    def __ne__(s, o): return F.not_equal(s, o)
    def __gt__(s, o): return F.greater(s, o)
    def __ge__(s, o): return F.greater_equal(s, o)
    def __lt__(s, o): return F.less(s, o)
    def __le__(s, o): return F.less_equal(s, o)
    def __and__(s, o): return F.logical_and(s, o) if s.dtype == BOOL else NotImplemented
    def __or__(s, o): return F.logical_or(s, o) if s.dtype == BOOL else NotImplemented
    def __xor__(s, o): return F.logical_xor(s, o) if s.dtype == BOOL else NotImplemented
    __hash__ = None

    # in-place family: one kernel, output aliasing input 0 (reference tensor.py:269-362)
    def _check_writeable(self):
        if not self.writeable:
            raise ValueError("output array is read-only")

    def _inplace(self, op, other):
        self._check_writeable()
        F.elementwise_into(op, self, self, other)
        return self

    def __iadd__(s, o): return s._inplace("ADD", o)
    def __isub__(s, o): return s._inplace("SUB", o)
    def __imul__(s, o): return s._inplace("MUL", o)
    def __itruediv__(s, o): return s._inplace("DIV", o)
    def __ifloordiv__(s, o): return s._inplace("FLOORDIV", o)
    def __imod__(s, o): return s._inplace("MOD", o)
    def __ipow__(s, o): return s._inplace("POW", o)

    def __imatmul__(s, o):
        s._check_writeable()
        r = F.matmul(s, o)
        if r.shape != s.shape:
            raise ValueError("inplace matrix multiplication requires the result to keep the shape")
        F.copy_into(s, r)
        return s

    # ------------------------------------------------------------------ indexing
    def __getitem__(self, key):
        return F.getitem(self, key)

    def __setitem__(self, key, value):
        self._check_writeable()
        F.setitem(self, key, value)


from . import functions as F  # noqa: E402  (cyclic by design: F builds DeviceArrays)
