"""The backend function table on the B200: every name the reference's `minidiff.backend` namespace
provides (the 114 attributes of minidiff/backend/numpy.py:14-206), with NumPy's
signatures, shape / broadcasting / dtype-promotion rules and error types, computing on device
memory through the C ABI (include/minidiff_b200.h).  NumPy itself is used here only for *metadata*
(dtype objects, `result_type`, host<->device staging of user data) -- never for arithmetic.
"""
from __future__ import annotations

import ctypes as C
import math
import operator

import numpy as np

from . import _lib
from ._lib import OP, RED, MdbArray, check, lib
from .device_array import BOOL, F32, F64, I64, DeviceArray, c_strides, dtype_code

_byref = C.byref
_pyscalar = (bool, int, float)
AxisError = np.exceptions.AxisError


# =============================================================================== plumbing
def asarray(x, dtype=None) -> DeviceArray:
    """tensor_constructor (np.array semantics, backend/numpy.py:15): host data -> device."""
    if isinstance(x, DeviceArray):
        return x if dtype is None or np.dtype(dtype) == x.dtype else astype(x, dtype)
    if isinstance(x, (list, tuple)) and _contains_device(x):
        if x and all(isinstance(e, DeviceArray) and e.shape == x[0].shape for e in x):
            r = stack(list(x))                       # a list of device arrays: stacked on the device
            return r if dtype is None or np.dtype(dtype) == r.dtype else astype(r, dtype)
        x = _to_host_nested(x)
    a = np.asarray(x) if dtype is None else np.asarray(x, dtype=dtype)
    return DeviceArray.from_numpy(a)


def _contains_device(seq) -> bool:
    for e in seq:
        if isinstance(e, DeviceArray) or (isinstance(e, (list, tuple)) and _contains_device(e)):
            return True
    return False


def _to_host_nested(seq):
    return [e.numpy() if isinstance(e, DeviceArray)
            else _to_host_nested(e) if isinstance(e, (list, tuple)) else e for e in seq]


def tensor_constructor(obj=(), dtype=None, copy=True, **_kw) -> DeviceArray:
    if isinstance(obj, DeviceArray):
        out = obj if dtype is None or np.dtype(dtype) == obj.dtype else astype(obj, dtype)
        return copy_(out) if copy and out is obj else out
    return asarray(obj, dtype)


def _operand(x):
    """DeviceArray | Python scalar (weak) | numpy scalar (strong) ; everything else is uploaded."""
    if isinstance(x, DeviceArray) or isinstance(x, _pyscalar) or isinstance(x, np.generic):
        return x
    return asarray(x)


def _dtype_token(x):
    if isinstance(x, DeviceArray):
        return x.dtype
    if isinstance(x, np.generic):
        return x.dtype
    return x  # Python scalar: weak under NEP 50


def result_dtype(*xs):
    first = xs[0]
    if len(xs) == 2 and first.__class__ is DeviceArray:       # the two commonest cases, decided at once
        second = xs[1]
        c = second.__class__
        if c is DeviceArray:
            if second.dtype is first.dtype:
                return first.dtype
        elif c is float:
            if first.dtype.kind == "f":
                return first.dtype
        elif c is int and first.dtype.kind in "iuf":
            return first.dtype
    if isinstance(first, DeviceArray):
        dt = first.dtype
        for x in xs[1:]:
            if isinstance(x, DeviceArray):
                if x.dtype != dt:
                    break
            elif isinstance(x, np.generic):   # NumPy scalars are strongly typed (np.float64 is a float!)
                break
            elif isinstance(x, bool) or (isinstance(x, int) and dt.kind in "iuf") or (
                    isinstance(x, float) and dt.kind == "f"):
                continue
            else:
                break
        else:
            return dt
    return np.result_type(*[_dtype_token(x) for x in xs])


def broadcast_shapes(shapes):
    if len(shapes) == 1:
        return shapes[0]
    if len(shapes) == 2 and shapes[0] == shapes[1]:
        return shapes[0]
    nd = max(len(s) for s in shapes)
    out = [1] * nd
    for s in shapes:
        off = nd - len(s)
        for i, e in enumerate(s):
            o = out[off + i]
            if e != 1:
                if o == 1:
                    out[off + i] = e
                elif o != e:
                    raise ValueError("operands could not be broadcast together with shapes "
                                     + " ".join(str(tuple(t)) for t in shapes))
    return tuple(out)


def _fill_imm(d: MdbArray, v):
    d.ptr = None
    d.ndim = 0
    if isinstance(v, np.generic):
        d.dtype = dtype_code(v.dtype)
        v = v.item()
    else:
        d.dtype = _lib.F64 if isinstance(v, float) else _lib.I64
    fv = float(v)
    d.imm = fv
    d.imm_i = int(v) if (not isinstance(v, float) or (math.isfinite(fv) and abs(fv) < 9e18)) else 0


def _launch_ew(op: str, out: DeviceArray, operands) -> DeviceArray:
    n = len(operands)
    descs = (MdbArray * n)()
    for i, o in enumerate(operands):
        if isinstance(o, DeviceArray):
            descs[i] = o.d
        else:
            _fill_imm(descs[i], o)
    check(lib.mdb_elementwise(OP[op], _byref(out.d), n, descs))
    return out


# ---- hot path of every elementwise backend function -------------------------------------------------
# What repeats from call to call is cached per (op, output dtype, operand shapes): the broadcast shape,
# the output strides and a byte image of the output descriptor.  A call then costs one descriptor copy,
# immediates written into three scratch descriptors, ONE ABI crossing (mdb_elementwise_new allocates the
# output and launches) and the DeviceArray that adopts the result.  Anything unusual (NumPy scalars,
# host data, all-scalar calls, > 3 operands) takes the general path below.
_PLANS: dict = {}
_IMM_DESC = [MdbArray() for _ in range(3)]
_IMM_REF = [_byref(d) for d in _IMM_DESC]
_F64_CODE, _I64_CODE = _lib.F64, _lib.I64
_new_call = lib.mdb_elementwise_new
_desc_from = MdbArray.from_buffer_copy


def _make_plan(key, op, out_dtype, shapes):
    shape = broadcast_shapes([s for s in shapes if s is not None])
    if len(shape) > _lib.MAX_DIMS:
        raise ValueError(f"at most {_lib.MAX_DIMS} dimensions are supported")
    strides = c_strides(shape)
    d = MdbArray()
    d.ptr = None
    d.dtype = dtype_code(out_dtype)
    d.ndim = len(shape)
    for i, (e, st) in enumerate(zip(shape, strides)):
        d.shape[i] = e
        d.strides[i] = st
    plan = (OP[op], shape, strides, math.prod(shape), bytes(d))
    if len(_PLANS) > 4096:
        _PLANS.clear()
    _PLANS[key] = plan
    return plan


def _imm_ref(slot, v):
    d = _IMM_DESC[slot]
    if v.__class__ is float:
        d.dtype = _F64_CODE
        d.imm = v
        d.imm_i = int(v) if (v == v and -9e18 < v < 9e18) else 0
    else:
        d.dtype = _I64_CODE
        d.imm = v
        d.imm_i = v
    return _IMM_REF[slot]


def _ew_into(op: str, out: DeviceArray, operands) -> bool:
    """`out = op(*operands)` into an existing array through the one-crossing entry point; operands
    are DeviceArrays or plain Python scalars that broadcast against out (the caller guarantees it).
    Returns False when an operand needs the general path."""
    n = len(operands)
    if n > 3:
        return False
    refs = [None, None, None]
    for i, o in enumerate(operands):
        c = o.__class__
        if c is DeviceArray:
            dd = o._desc
            refs[i] = _byref(dd if dd is not None else o.d)
        elif (c is float or c is int) and -9.2e18 < o < 9.2e18:
            refs[i] = _imm_ref(i, o)
        else:
            return False
    dd = out._desc
    rc = _new_call(OP[op], dd if dd is not None else out.d, n, refs[0], refs[1], refs[2])
    if rc:
        check(rc)
    return True


def _ew(op: str, out_dtype, *operands) -> DeviceArray:
    n = len(operands)
    if n <= 3:
        shapes = []
        fast = False
        for o in operands:
            c = o.__class__
            if c is DeviceArray:
                shapes.append(o.shape)
                fast = True
            elif (c is float or c is int) and -9.2e18 < o < 9.2e18 or c is bool:
                shapes.append(None)
            else:
                fast = False
                break
        if fast:
            key = (op, out_dtype, *shapes)
            plan = _PLANS.get(key)
            if plan is None:
                plan = _make_plan(key, op, np.dtype(out_dtype), shapes)
            opid, shape, strides, size, image = plan
            d = _desc_from(image)
            refs = [None, None, None]
            for i, o in enumerate(operands):
                if o.__class__ is DeviceArray:
                    dd = o._desc
                    refs[i] = _byref(dd if dd is not None else o.d)
                else:
                    refs[i] = _imm_ref(i, int(o) if o.__class__ is bool else o)
            rc = _new_call(opid, d, n, refs[0], refs[1], refs[2])
            if rc:
                check(rc)
            return DeviceArray._adopt(d, shape, strides, out_dtype if isinstance(out_dtype, np.dtype) else np.dtype(out_dtype), size)
    operands = [_operand(o) for o in operands]
    shape = broadcast_shapes([o.shape for o in operands if isinstance(o, DeviceArray)] or [()])
    if not any(isinstance(o, DeviceArray) for o in operands):
        operands[0] = asarray(operands[0])  # all-scalar call: NumPy returns a 0-d result
        shape = ()
    return _launch_ew(op, DeviceArray.empty(shape, out_dtype), operands)


def _byte_extent(a: DeviceArray):
    lo = hi = 0
    for e, st in zip(a.shape, a.estrides):
        if e == 0:
            return a.ptr, a.ptr
        if st < 0:
            lo += (e - 1) * st
        else:
            hi += (e - 1) * st
    return a.ptr + lo * a.itemsize, a.ptr + (hi + 1) * a.itemsize


def _clashes(out: DeviceArray, x) -> bool:
    """True when writing `out` elementwise while reading `x` is a hazard: same allocation, byte ranges
    intersect, and x is not the very same view (exact aliasing is what the in-place kernels support).
    NumPy detects such overlap and buffers the operand (`a += a.T`, `a[1:] = a[:-1]`, `x -= x[::-1]`)."""
    if not isinstance(x, DeviceArray) or x._st is not out._st:
        return False
    if x.ptr == out.ptr and x.shape == out.shape and x.estrides == out.estrides:
        return False
    lo_o, hi_o = _byte_extent(out)
    lo_x, hi_x = _byte_extent(x)
    return lo_o < hi_x and lo_x < hi_o


def elementwise_into(op: str, out: DeviceArray, *operands) -> DeviceArray:
    """`out = op(*operands)` written in place (out may be operands[0]): the `+=` family."""
    operands = [copy_(o) if _clashes(out, o) else o for o in (_operand(o) for o in operands)]
    rdt = result_dtype(*operands)
    if rdt != out.dtype and not np.can_cast(rdt, out.dtype, casting="same_kind"):
        raise TypeError(f"Cannot cast ufunc '{op.lower()}' output from {rdt!r} to {out.dtype!r} "
                        "with casting rule 'same_kind'")
    shape = broadcast_shapes([o.shape for o in operands if isinstance(o, DeviceArray)] or [()])
    if shape != out.shape and broadcast_shapes([shape, out.shape]) != out.shape:
        raise ValueError(f"non-broadcastable output operand with shape {out.shape} doesn't match "
                         f"the broadcast shape {shape}")
    if _ew_into(op, out, operands):
        return out
    return _launch_ew(op, out, operands)


def copy_into(dst: DeviceArray, src) -> None:
    """dst[...] = src with broadcasting + dtype cast (basic `__setitem__`, astype, materialising)."""
    src = _operand(src)
    if isinstance(src, DeviceArray):
        if _clashes(dst, src):                      # overlapping views of one allocation: buffer the source
            tmp = DeviceArray.empty(src.shape, src.dtype)
            check(lib.mdb_copy(_byref(tmp.d), _byref(src.d)))
            src = tmp
        check(lib.mdb_copy(_byref(dst.d), _byref(src.d)))
    else:
        d = MdbArray()
        _fill_imm(d, src)
        check(lib.mdb_copy(_byref(dst.d), _byref(d)))


def _float_dtype_for(x):
    dt = x.dtype if isinstance(x, (DeviceArray, np.generic)) else np.result_type(x)
    if dt.kind == "f":
        return F32 if dt == np.float16 else dt
    if dt.kind in "biu" and dt.itemsize <= 2:
        return F32  # NumPy would give float16/float32; fp16 storage is not supported here
    return F64


def _unary_float(op):
    def f(x, **_kw):
        x = asarray(x) if not isinstance(x, DeviceArray) else x
        return _ew(op, _float_dtype_for(x), x)

    f.__name__ = op.lower()
    return f


def _unary_same(op):
    def f(x, **_kw):
        x = asarray(x) if not isinstance(x, DeviceArray) else x
        return _ew(op, x.dtype, x)

    f.__name__ = op.lower()
    return f


def _binary_arith(op, name):
    def f(x, y, **_kw):
        cx, cy = x.__class__, y.__class__
        if not (cx is DeviceArray or cx is float or cx is int):
            x = _operand(x)
        if not (cy is DeviceArray or cy is float or cy is int):
            y = _operand(y)
        return _ew(op, result_dtype(x, y), x, y)

    f.__name__ = name
    return f


def _binary_pred(op, name):
    def f(x, y, **_kw):
        return _ew(op, BOOL, x, y)

    f.__name__ = name
    return f


# =============================================================================== elementwise table
sin, cos, tan = _unary_float("SIN"), _unary_float("COS"), _unary_float("TAN")
sinh, cosh, tanh = _unary_float("SINH"), _unary_float("COSH"), _unary_float("TANH")
exp, log = _unary_float("EXP"), _unary_float("LOG")
ceil, floor, sign = _unary_same("CEIL"), _unary_same("FLOOR"), _unary_same("SIGN")
absolute, negative = _unary_same("ABS"), _unary_same("NEG")
add = _binary_arith("ADD", "add")
subtract = _binary_arith("SUB", "subtract")
multiply = _binary_arith("MUL", "multiply")
mod = _binary_arith("MOD", "mod")
floor_divide = _binary_arith("FLOORDIV", "floor_divide")
equal, not_equal = _binary_pred("EQ", "equal"), _binary_pred("NE", "not_equal")
greater, greater_equal = _binary_pred("GT", "greater"), _binary_pred("GE", "greater_equal")
less, less_equal = _binary_pred("LT", "less"), _binary_pred("LE", "less_equal")
logical_and = _binary_pred("AND", "logical_and")
logical_or = _binary_pred("OR", "logical_or")
logical_xor = _binary_pred("XOR", "logical_xor")


def logical_not(x, **_kw):
    return _ew("LOGICAL_NOT", BOOL, x)


def invert(x, **_kw):
    x = asarray(x)
    if x.dtype == BOOL:
        return _ew("LOGICAL_NOT", BOOL, x)
    if x.dtype.kind not in "iu":
        raise TypeError("ufunc 'invert' not supported for the input types")
    return _ew("INVERT", x.dtype, x)


def true_divide(x, y, **_kw):
    x, y = _operand(x), _operand(y)
    dt = result_dtype(x, y)
    if dt.kind in "biu":
        dt = F64
    return _ew("DIV", dt, x, y)


def power(x, y, **_kw):
    x, y = _operand(x), _operand(y)
    dt = result_dtype(x, y)
    if dt.kind in "iu" and isinstance(y, int) and not isinstance(y, bool) and y < 0:
        raise ValueError("Integers to negative integer powers are not allowed.")
    if dt == BOOL:
        dt = np.dtype(np.int8)
    return _ew("POW", dt, x, y)


def where(condition, x=None, y=None):
    if x is None or y is None:
        raise NotImplementedError("where(condition) without x, y is not supported on device")
    x, y = _operand(x), _operand(y)
    return _ew("WHERE", result_dtype(x, y), _operand(condition), x, y)


def clip(a, a_min=None, a_max=None, **_kw):
    a = asarray(a)
    if a_min is None and a_max is None:
        raise ValueError("One of max or min must be given")
    lo = -math.inf if a_min is None else _operand(a_min)
    hi = math.inf if a_max is None else _operand(a_max)
    parts = [a] + [v for v, given in ((lo, a_min is not None), (hi, a_max is not None)) if given]
    dt = result_dtype(*parts)
    if dt.kind in "iu":  # +-inf immediates are not representable: use the dtype's extremes
        info = np.iinfo(dt)
        lo = info.min if a_min is None else lo
        hi = info.max if a_max is None else hi
    return _ew("CLIP", dt, a, lo, hi)


def astype(a, dtype, copy=True, **_kw):
    dtype = np.dtype(dtype)
    dtype_code(dtype)
    if not copy and dtype == a.dtype:
        return a
    out = DeviceArray.empty(a.shape, dtype)
    copy_into(out, a)
    return out


def copy_(a, **_kw):
    a = asarray(a)
    out = DeviceArray.empty(a.shape, a.dtype)
    copy_into(out, a)
    return out


# =============================================================================== views (no kernels)
def _norm_axis(ax, nd):
    ax = operator.index(ax)
    if not -nd <= ax < nd:
        raise AxisError(ax, nd)
    return ax % nd if nd else 0


def _norm_axes(axis, nd, allow_dup=False):
    if axis is None:
        return tuple(range(nd))
    if isinstance(axis, DeviceArray):
        axis = axis.numpy().tolist()
    if isinstance(axis, (int, np.integer)):
        axis = (axis,)
    out = tuple(_norm_axis(a, nd) for a in axis)
    if not allow_dup and len(set(out)) != len(out):
        raise ValueError("duplicate value in 'axis'")
    return out


def transpose(a, axes=None):
    a = asarray(a)
    if axes is None:
        return a.T
    if isinstance(axes, DeviceArray):
        axes = axes.numpy().tolist()
    axes = [int(x.item()) if isinstance(x, DeviceArray) else operator.index(x) for x in axes]
    if len(axes) != a.ndim:
        raise ValueError("axes don't match array")
    axes = [_norm_axis(x, a.ndim) for x in axes]
    if sorted(axes) != list(range(a.ndim)):
        raise ValueError("repeated axis in transpose")
    return a.view([a.shape[i] for i in axes], [a.estrides[i] for i in axes])


def swapaxes(a, axis1, axis2):
    a = asarray(a)
    i, j = _norm_axis(axis1, a.ndim), _norm_axis(axis2, a.ndim)
    perm = list(range(a.ndim))
    perm[i], perm[j] = perm[j], perm[i]
    return transpose(a, perm)


def broadcast_to(a, shape, **_kw):
    a = asarray(a)
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    if len(shape) < a.ndim:
        raise ValueError("input operand has more dimensions than allowed by the axis remapping")
    lead = len(shape) - a.ndim
    st = [0] * lead
    for e, s, t in zip(a.shape, a.estrides, shape[lead:]):
        if e == t:
            st.append(s)
        elif e == 1:
            st.append(0)
        else:
            raise ValueError(f"operands could not be broadcast together with remapped shapes "
                             f"[original->remapped]: {a.shape} and requested shape {shape}")
    return a.view(shape, st, writeable=False)  # NumPy's broadcast_to result is read-only


def expand_dims(a, axis):
    a = asarray(a)
    if isinstance(axis, (int, np.integer)):
        axis = (axis,)
    nd = a.ndim + len(axis)
    axis = _norm_axes(axis, nd)
    shape, st, it = [], [], iter(zip(a.shape, a.estrides))
    for i in range(nd):
        if i in axis:
            shape.append(1); st.append(0)
        else:
            e, s = next(it)
            shape.append(e); st.append(s)
    return a.view(shape, st)


def squeeze(a, axis=None):
    a = asarray(a)
    if axis is None:
        drop = [i for i, e in enumerate(a.shape) if e == 1]
    else:
        drop = _norm_axes(axis, a.ndim)
        for i in drop:
            if a.shape[i] != 1:
                raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
    keep = [i for i in range(a.ndim) if i not in drop]
    return a.view([a.shape[i] for i in keep], [a.estrides[i] for i in keep])


def flip(a, axis=None):
    a = asarray(a)
    axes = _norm_axes(axis, a.ndim)
    ptr, st = a.ptr, list(a.estrides)
    for i in axes:
        if a.shape[i] > 1:
            ptr += (a.shape[i] - 1) * st[i] * a.itemsize
            st[i] = -st[i]
    return a.view(a.shape, st, ptr=ptr)


def _atleast(a, n):
    a = asarray(a)
    if a.ndim >= n:
        return a
    if n == 1:
        return a.view((1,), (0,))
    if n == 2:
        return a.view((1,) * (2 - a.ndim) + a.shape, (0,) * (2 - a.ndim) + a.estrides)
    if a.ndim == 0:
        return a.view((1, 1, 1), (0, 0, 0))
    if a.ndim == 1:
        return a.view((1, a.shape[0], 1), (0, a.estrides[0], 0))
    return a.view(a.shape + (1,), a.estrides + (0,))


def atleast_1d(a): return _atleast(a, 1)
def atleast_2d(a): return _atleast(a, 2)
def atleast_3d(a): return _atleast(a, 3)


def _reshape_strides(shape, estrides, newshape):
    """NumPy's no-copy reshape rule (_attempt_nocopy_reshape): new strides or None."""
    old = [(e, s) for e, s in zip(shape, estrides) if e != 1]
    if not old:
        return c_strides(newshape)
    new_st = [0] * len(newshape)
    oi = oj = ni = nj = 0
    on, nn = len(old), len(newshape)
    ni, nj, oi, oj = 0, 1, 0, 1
    while ni < nn and oi < on:
        np_, op_ = newshape[ni], old[oi][0]
        while np_ != op_:
            if np_ < op_:
                np_ *= newshape[nj]; nj += 1
            else:
                op_ *= old[oj][0]; oj += 1
        for k in range(oi, oj - 1):
            if old[k][1] != old[k + 1][0] * old[k + 1][1]:
                return None
        new_st[nj - 1] = old[oj - 1][1]
        for k in range(nj - 1, ni, -1):
            new_st[k - 1] = new_st[k] * newshape[k]
        ni, nj = nj, nj + 1
        oi, oj = oj, oj + 1
    last = new_st[ni - 1] if ni >= 1 else 1
    for k in range(ni, nn):
        new_st[k] = last
    return tuple(new_st)


def reshape(a, shape=None, order="C", newshape=None, **_kw):
    a = asarray(a)
    if shape is None:
        shape = newshape
    if isinstance(shape, DeviceArray):
        shape = shape.numpy().tolist()
    shape = [int(shape)] if isinstance(shape, (int, np.integer)) else [int(s) for s in shape]
    if order == "F":
        return reshape(a.T, shape[::-1]).T
    if order not in ("C", "A", None):
        raise ValueError("order must be 'C' or 'F'")
    if shape.count(-1) > 1:
        raise ValueError("can only specify one unknown dimension")
    if -1 in shape:
        known = -math.prod(shape)
        if known == 0 or a.size % known:
            raise ValueError(f"cannot reshape array of size {a.size} into shape {tuple(shape)}")
        shape[shape.index(-1)] = a.size // known
    if math.prod(shape) != a.size:
        raise ValueError(f"cannot reshape array of size {a.size} into shape {tuple(shape)}")
    if len(shape) > _lib.MAX_DIMS:
        raise ValueError(f"at most {_lib.MAX_DIMS} dimensions are supported")
    if a.size == 0:
        return a.view(shape, c_strides(shape))
    st = _reshape_strides(a.shape, a.estrides, shape)
    if st is None:
        a = copy_(a)
        st = c_strides(shape)
    return a.view(shape, st)


def ravel(a, order="C"):
    a = asarray(a)
    return reshape(a, (a.size,), order=order)


def flatten(a, order="C"):
    a = asarray(a)
    if order == "F":
        a = a.T
    return reshape(copy_(a), (a.size,))


# =============================================================================== reductions
def _reduce(red, a, axis, keepdims, out_dtype):
    a = asarray(a)
    nd = a.ndim
    axes = _norm_axes(axis, nd)
    kshape = tuple(1 if i in axes else e for i, e in enumerate(a.shape))
    out = DeviceArray.empty(kshape, out_dtype)
    mask = 0
    for i in axes:
        mask |= 1 << i
    if red in ("MAX", "MIN", "ARGMAX", "ARGMIN") and a.size == 0:
        raise ValueError("zero-size array to reduction operation which has no identity")
    check(lib.mdb_reduce(RED[red], _byref(out.d), _byref(a.d), mask))
    if keepdims:
        return out
    keep = [i for i in range(nd) if i not in axes]
    return out.view([kshape[i] for i in keep], [out.estrides[i] for i in keep])


def sum_(a, axis=None, keepdims=False, dtype=None, **_kw):
    a = asarray(a)
    dt = np.dtype(dtype) if dtype is not None else (
        a.dtype if a.dtype.kind == "f" else np.dtype(np.uint64) if a.dtype == np.uint64 else I64)
    return _reduce("SUM", a, axis, keepdims, dt)


def prod_(a, axis=None, keepdims=False, dtype=None, **_kw):
    a = asarray(a)
    dt = np.dtype(dtype) if dtype is not None else (a.dtype if a.dtype.kind == "f" else I64)
    return _reduce("PROD", a, axis, keepdims, dt)


def mean(a, axis=None, keepdims=False, dtype=None, **_kw):
    a = asarray(a)
    dt = np.dtype(dtype) if dtype is not None else (a.dtype if a.dtype.kind == "f" else F64)
    return _reduce("MEAN", a, axis, keepdims, dt)


def max_(a, axis=None, keepdims=False, **_kw):
    a = asarray(a)
    return _reduce("MAX", a, axis, keepdims, a.dtype)


def min_(a, axis=None, keepdims=False, **_kw):
    a = asarray(a)
    return _reduce("MIN", a, axis, keepdims, a.dtype)


def any_(a, axis=None, keepdims=False, **_kw):
    return _reduce("ANY", a, axis, keepdims, BOOL)


def all_(a, axis=None, keepdims=False, **_kw):
    return _reduce("ALL", a, axis, keepdims, BOOL)


def _argreduce(red, a, axis, keepdims):
    a = asarray(a)
    if axis is None:
        flat = ravel(a)
        r = _reduce(red, flat, 0, False, I64)
        return r.view((1,) * a.ndim, (0,) * a.ndim) if keepdims else r
    if not isinstance(axis, (int, np.integer)):
        raise TypeError(f"'{type(axis).__name__}' object cannot be interpreted as an integer")
    return _reduce(red, a, axis, keepdims, I64)


def argmax(a, axis=None, keepdims=False, **_kw):
    return _argreduce("ARGMAX", a, axis, keepdims)


def argmin(a, axis=None, keepdims=False, **_kw):
    return _argreduce("ARGMIN", a, axis, keepdims)


def std(a, axis=None, keepdims=False, ddof=0, **_kw):
    a = asarray(a)
    if isinstance(axis, DeviceArray):
        axis = tuple(axis.numpy().tolist())
    fdt = a.dtype if a.dtype.kind == "f" else F64
    axes = _norm_axes(axis, a.ndim)
    n = math.prod(a.shape[i] for i in axes)
    mu = mean(a, axis=axes, keepdims=True)
    dev = _ew("SUB", fdt, a, mu)
    sq = _ew("SQUARE", fdt, dev)
    var = true_divide(sum_(sq, axis=axes, keepdims=keepdims), max(n - ddof, 0))
    return _ew("SQRT", fdt, var)


# =============================================================================== contractions
def _gemm(a: DeviceArray, b: DeviceArray, out=None, accumulate=False) -> DeviceArray:
    if a.shape[1] != b.shape[0]:
        raise ValueError(
            f"matmul: Input operand 1 has a mismatch in its core dimension 0, with gufunc signature "
            f"(n?,k),(k,m?)->(n?,m?) (size {b.shape[0]} is different from {a.shape[1]})")
    if out is None:
        out = DeviceArray.empty((a.shape[0], b.shape[1]), a.dtype)
    check(lib.mdb_gemm(_byref(out.d), _byref(a.d), _byref(b.d), 1 if accumulate else 0))
    return out


def _gemm_fused(a: DeviceArray, b: DeviceArray, bias=None, relu=False, mask_src=None, out=None,
                accumulate=False):
    """out (+)= mask(relu?(a @ b + bias?)) in ONE launch of the CTA-pair tcgen05 kernel (C ABI
    mdb_gemm_fused).  Returns None when the problem cannot run there (small / unaligned shapes):
    the caller then issues the unfused chain of backend calls."""
    if a.shape[1] != b.shape[0]:
        raise ValueError(
            f"matmul: Input operand 1 has a mismatch in its core dimension 0, with gufunc signature "
            f"(n?,k),(k,m?)->(n?,m?) (size {b.shape[0]} is different from {a.shape[1]})")
    m, n = a.shape[0], b.shape[1]
    if m <= 128 or n <= 128 or a.shape[1] < 32 or m * n * a.shape[1] < (1 << 21):
        return None
    if bias is not None and not (bias.dtype == F32 and bias.shape == (n,) and bias.estrides == (1,)):
        return None
    if mask_src is not None and not (mask_src.dtype == F32 and mask_src.shape == (m, n) and mask_src.estrides[1] == 1):
        return None
    if out is None:
        out = DeviceArray.empty((m, n), F32)
    rc = lib.mdb_gemm_fused(_byref(out.d), _byref(a.d), _byref(b.d), 1 if accumulate else 0,
                            _byref(bias.d) if bias is not None else None, 1 if relu else 0,
                            _byref(mask_src.d) if mask_src is not None else None)
    if rc == _lib.ENOTSUP:
        return None
    check(rc)
    return out


def _gemm_batched(xv: DeviceArray, yv: DeviceArray) -> DeviceArray:
    """np.matmul for stacked and/or float64 operands: every matrix of the (broadcast) batch in ONE
    launch of the CUDA-core kernel (C ABI mdb_gemm_batched); O(M*N) memory per matrix."""
    if xv.shape[-1] != yv.shape[-2]:
        raise ValueError(
            f"matmul: Input operand 1 has a mismatch in its core dimension 0, with gufunc signature "
            f"(n?,k),(k,m?)->(n?,m?) (size {yv.shape[-2]} is different from {xv.shape[-1]})")
    batch = broadcast_shapes([xv.shape[:-2], yv.shape[:-2]])
    nb = len(batch)
    xa = xv.view((1,) * (nb - (xv.ndim - 2)) + xv.shape, (0,) * (nb - (xv.ndim - 2)) + xv.estrides)
    ya = yv.view((1,) * (nb - (yv.ndim - 2)) + yv.shape, (0,) * (nb - (yv.ndim - 2)) + yv.estrides)
    out = DeviceArray.empty(batch + (xv.shape[-2], yv.shape[-1]), xv.dtype)
    if out.size:
        if xv.shape[-1] == 0:
            copy_into(out, 0)
        else:
            check(lib.mdb_gemm_batched(_byref(out.d), _byref(xa.d), _byref(ya.d)))
    return out


def matmul(x, y, **_kw):
    x, y = asarray(x), asarray(y)
    if x.ndim == 0 or y.ndim == 0:
        raise ValueError("matmul: Input operand does not have enough dimensions")
    dt = result_dtype(x, y)
    if dt.kind != "f" or dt == np.float16:
        dt = F64 if dt.kind in "iu" and dt.itemsize >= 4 else F32
    rdt = result_dtype(x, y)
    x = x if x.dtype == dt else astype(x, dt)
    y = y if y.dtype == dt else astype(y, dt)
    xv = x.view((1,) + x.shape, (0,) + x.estrides) if x.ndim == 1 else x
    yv = y.view(y.shape + (1,), y.estrides + (0,)) if y.ndim == 1 else y
    if dt == F32 and xv.ndim == 2 and yv.ndim == 2:
        r = _gemm(xv, yv)                 # tcgen05 3xTF32 (CUDA-core kernel for small / odd shapes)
    elif dt == F32 and xv.shape[-2] * yv.shape[-1] * xv.shape[-1] >= (1 << 27) and len(
            broadcast_shapes([xv.shape[:-2], yv.shape[:-2]])) <= 2 and math.prod(
            broadcast_shapes([xv.shape[:-2], yv.shape[:-2]])) <= 64:
        # a few LARGE stacked matrices: one tensor-core GEMM each beats the batched CUDA-core kernel
        if xv.shape[-1] != yv.shape[-2]:
            raise ValueError("matmul: Input operand 1 has a mismatch in its core dimension 0")
        batch = broadcast_shapes([xv.shape[:-2], yv.shape[:-2]])
        xb = broadcast_to(xv, batch + xv.shape[-2:])
        yb = broadcast_to(yv, batch + yv.shape[-2:])
        r = DeviceArray.empty(batch + (xv.shape[-2], yv.shape[-1]), dt)
        for idx in np.ndindex(*batch):
            _gemm(getitem(xb, idx), getitem(yb, idx), out=getitem(r, idx))
    else:
        r = _gemm_batched(xv, yv)         # float64 / integer operands and stacked matrices: one launch
    if x.ndim == 1:
        r = squeeze(r, axis=-2)
    if y.ndim == 1:
        r = squeeze(r, axis=-1)
    return r if rdt == dt else astype(r, rdt)


def dot(a, b, **_kw):
    a, b = _operand(a), _operand(b)
    if not isinstance(a, DeviceArray) or not isinstance(b, DeviceArray) or a.ndim == 0 or b.ndim == 0:
        return multiply(a, b)
    if a.ndim == 1 and b.ndim == 1:
        if a.shape != b.shape:
            raise ValueError(f"shapes {a.shape} and {b.shape} not aligned")
        return sum_(multiply(a, b))
    if b.ndim <= 2:
        return matmul(a, b)
    return tensordot(a, b, axes=((a.ndim - 1,), (b.ndim - 2,)))


def tensordot(a, b, axes=2):
    a, b = asarray(a), asarray(b)
    if isinstance(axes, (int, np.integer)):
        ax_a, ax_b = list(range(a.ndim - axes, a.ndim)), list(range(axes))
    else:
        ax_a, ax_b = axes
        ax_a = [ax_a] if isinstance(ax_a, (int, np.integer)) else list(ax_a)
        ax_b = [ax_b] if isinstance(ax_b, (int, np.integer)) else list(ax_b)
    ax_a = [_norm_axis(i, a.ndim) for i in ax_a]
    ax_b = [_norm_axis(i, b.ndim) for i in ax_b]
    if len(ax_a) != len(ax_b) or any(a.shape[i] != b.shape[j] for i, j in zip(ax_a, ax_b)):
        raise ValueError("shape-mismatch for sum")
    free_a = [i for i in range(a.ndim) if i not in ax_a]
    free_b = [i for i in range(b.ndim) if i not in ax_b]
    k = math.prod(a.shape[i] for i in ax_a)
    at = reshape(transpose(a, free_a + ax_a), (-1, k) if k else (math.prod(a.shape[i] for i in free_a), 0))
    bt = reshape(transpose(b, ax_b + free_b), (k, -1) if k else (0, math.prod(b.shape[i] for i in free_b)))
    r = matmul(at, bt)
    return reshape(r, [a.shape[i] for i in free_a] + [b.shape[i] for i in free_b])


# =============================================================================== indexing
def _is_index_array(k):
    return isinstance(k, (DeviceArray, np.ndarray, list))


def _basic_view(a: DeviceArray, key):
    if not isinstance(key, tuple):
        key = (key,)
    n_specified = sum(1 for k in key if k is not None and k is not Ellipsis)
    if n_specified > a.ndim:
        raise IndexError(f"too many indices for array: array is {a.ndim}-dimensional, but "
                         f"{n_specified} were indexed")
    if sum(1 for k in key if k is Ellipsis) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    ptr, shape, st, dim = a.ptr, [], [], 0
    for k in key:
        if k is None:
            shape.append(1); st.append(0)
        elif k is Ellipsis:
            for _ in range(a.ndim - n_specified):
                shape.append(a.shape[dim]); st.append(a.estrides[dim]); dim += 1
        elif isinstance(k, slice):
            start, stop, step = k.indices(a.shape[dim])
            n = len(range(start, stop, step))
            ptr += start * a.estrides[dim] * a.itemsize if n else 0
            shape.append(n); st.append(a.estrides[dim] * step); dim += 1
        else:
            i = operator.index(k)
            e = a.shape[dim]
            if not -e <= i < e:
                raise IndexError(f"index {i} is out of bounds for axis {dim} with size {e}")
            ptr += (i % e) * a.estrides[dim] * a.itemsize
            dim += 1
    while dim < a.ndim:
        shape.append(a.shape[dim]); st.append(a.estrides[dim]); dim += 1
    return a.view(shape, st, ptr=ptr)


def _advanced_plan(a: DeviceArray, key):
    """Integer-array keys on the leading axes (optionally followed by full slices).  Returns
    (element offsets int64 [n], index result shape, trailing shape, trailing strides)."""
    if not isinstance(key, tuple):
        key = (key,)
    idx, rest = [], []
    for k in key:
        if rest or (isinstance(k, slice) and k == slice(None)):
            if not (isinstance(k, slice) and k == slice(None)):
                raise NotImplementedError("advanced indexing is supported on leading axes only")
            rest.append(k)
        elif _is_index_array(k) or isinstance(k, (int, np.integer)):
            idx.append(k)
        else:
            raise NotImplementedError(f"unsupported mixed index {key!r}")
    if len(idx) > a.ndim:
        raise IndexError("too many indices for array")
    arrs = []
    for k in idx:
        k = asarray(k) if not isinstance(k, (int, np.integer)) else k
        if isinstance(k, DeviceArray) and k.dtype == BOOL:
            raise NotImplementedError("boolean masks are handled separately")
        if isinstance(k, DeviceArray) and k.dtype.kind not in "iu":
            raise IndexError("arrays used as indices must be of integer (or boolean) type")
        arrs.append(k)
    bshape = broadcast_shapes([k.shape for k in arrs if isinstance(k, DeviceArray)] or [()])
    # element offsets, one fused + VALIDATED launch per index array (mdb_index_offsets: wraps negative
    # indices, raises IndexError for anything outside [-extent, extent) like NumPy does)
    off = DeviceArray.empty(bshape, I64)
    const, first = 0, True
    for axis, k in enumerate(arrs):
        e = a.shape[axis]
        if isinstance(k, DeviceArray):
            kb = k.view((1,) * (len(bshape) - k.ndim) + k.shape, (0,) * (len(bshape) - k.ndim) + k.estrides)
            check(lib.mdb_index_offsets(_byref(off.d), _byref(kb.d), e, a.estrides[axis], 0 if first else 1))
            first = False
        else:
            i = operator.index(k)
            if not -e <= i < e:
                raise IndexError(f"index {i} is out of bounds for axis {axis} with size {e}")
            const += (i % e) * a.estrides[axis]
    if first:
        copy_into(off, const)
    elif const:
        elementwise_into("ADD", off, off, const)
    off = reshape(off, (-1,))
    k = len(arrs)
    return off, bshape, a.shape[k:], a.estrides[k:]


def _rows_view(a: DeviceArray, tshape, tstrides) -> MdbArray:
    d = MdbArray()
    d.ptr = a.ptr
    d.dtype = dtype_code(a.dtype)
    d.ndim = 1 + len(tshape)
    d.shape[0] = 1 << 62          # MDB_ROWS_ARE_OFFSETS: the index vector holds validated element offsets
    d.strides[0] = 1
    for i, (e, s) in enumerate(zip(tshape, tstrides)):
        d.shape[i + 1] = e
        d.strides[i + 1] = s
    return d


def _mask_to_indices(mask: DeviceArray) -> DeviceArray:
    """Flat positions of the True entries of `mask`, in order: device stream compaction (ballot + scan,
    C ABI mdb_nonzero).  Only the COUNT comes back to the host (8 bytes) -- the result size is
    data-dependent, so like every GPU array library this synchronises once."""
    m = mask if mask.dtype == BOOL else not_equal(mask, 0)
    m = m if m.is_c_contiguous() else copy_(m)
    buf = DeviceArray.empty((m.size,), I64)
    n = C.c_int64(0)
    check(lib.mdb_nonzero(_byref(m.d), _byref(buf.d), _byref(n)))
    return buf.view((n.value,), (1,))


def getitem(a, key):
    a = asarray(a)
    if isinstance(key, DeviceArray) and key.dtype == BOOL or isinstance(key, np.ndarray) and key.dtype == np.bool_:
        key = asarray(key)
        if key.shape != a.shape[:key.ndim]:
            raise IndexError("boolean index did not match indexed array")
        idx = _mask_to_indices(key)
        flat = reshape(a, (math.prod(key.shape),) + a.shape[key.ndim:])
        return getitem(flat, idx)
    keys = key if isinstance(key, tuple) else (key,)
    if not any(_is_index_array(k) for k in keys):
        return _basic_view(a, key)
    off, bshape, tshape, tstrides = _advanced_plan(a, key)
    n = off.shape[0]
    out = DeviceArray.empty((n,) + tuple(tshape), a.dtype)
    src = _rows_view(a, tshape, tstrides)
    check(lib.mdb_gather_rows(_byref(out.d), _byref(src), _byref(off.d)))
    return reshape(out, tuple(bshape) + tuple(tshape))


def _scatter(a: DeviceArray, key, value, add_mode: bool):
    off, bshape, tshape, tstrides = _advanced_plan(a, key)
    n = off.shape[0]
    value = _operand(value)
    full = (n,) + tuple(tshape)
    if isinstance(value, DeviceArray):
        v = value if value.dtype == a.dtype else astype(value, a.dtype)
        v = broadcast_to(v, tuple(bshape) + tuple(tshape))
        v = reshape(v, full) if _reshape_strides(v.shape, v.estrides, list(full)) is not None \
            else reshape(copy_(v), full)
    else:
        v = DeviceArray.empty((1,) * len(full), a.dtype)
        copy_into(v, value)
        v = broadcast_to(v, full)
    dst = _rows_view(a, tshape, tstrides)
    check(lib.mdb_scatter_rows(_byref(dst), _byref(v.d), _byref(off.d), 1 if add_mode else 0))


def setitem(a: DeviceArray, key, value):
    if isinstance(key, DeviceArray) and key.dtype == BOOL:
        idx = _mask_to_indices(key)
        flat = reshape(a, (math.prod(key.shape),) + a.shape[key.ndim:])
        if flat._st is not a._st:
            raise NotImplementedError("boolean-mask assignment needs a contiguous target")
        return _scatter(flat, idx, value, False)
    keys = key if isinstance(key, tuple) else (key,)
    if not any(_is_index_array(k) for k in keys):
        return copy_into(_basic_view(a, key), value)
    return _scatter(a, key, value, False)


def index_add(a, indices, b=None):
    """np.add.at (backend/numpy.py:105): unbuffered a[indices] += b, duplicates accumulate."""
    if not a.writeable:
        raise ValueError("output array is read-only")
    keys = indices if isinstance(indices, tuple) else (indices,)
    if any(isinstance(k, DeviceArray) and k.dtype == BOOL for k in keys):
        sel = getitem(a, indices)
        return setitem(a, indices, add(sel, b))
    if not any(_is_index_array(k) for k in keys):
        v = _basic_view(a, indices)
        elementwise_into("ADD", v, v, b)
        return None
    _scatter(a, indices, b, True)
    return None


def _along_axis_offsets(arr: DeviceArray, indices: DeviceArray, axis: int) -> DeviceArray:
    if indices.ndim != arr.ndim:
        raise ValueError("`indices` and `arr` must have the same number of dimensions")
    if indices.dtype.kind not in "iu":
        raise IndexError("`indices` must be an integer array")
    off = DeviceArray.empty(indices.shape, I64)
    check(lib.mdb_index_offsets(_byref(off.d), _byref(indices.d), arr.shape[axis], arr.estrides[axis], 0))
    for d in range(arr.ndim):
        if d == axis or indices.shape[d] == 1 and arr.shape[d] == 1:
            continue
        shp = [1] * arr.ndim
        shp[d] = indices.shape[d]
        grid = reshape(arange(indices.shape[d]), shp)
        elementwise_into("ADD", off, off, multiply(grid, arr.estrides[d]))
    return reshape(off, (-1,))


def take_along_axis(arr, indices, axis=None):
    arr, indices = asarray(arr), asarray(indices)
    if axis is None:
        arr, axis = ravel(arr), 0
    axis = _norm_axis(axis, arr.ndim)
    off = _along_axis_offsets(arr, indices, axis)
    out = DeviceArray.empty((off.shape[0],), arr.dtype)
    check(lib.mdb_gather_rows(_byref(out.d), _byref(_rows_view(arr, (), ())), _byref(off.d)))
    return reshape(out, indices.shape)


def put_along_axis(arr, indices, values, axis):
    arr, indices = asarray(arr), asarray(indices)
    if not arr.writeable:
        raise ValueError("output array is read-only")
    if axis is None:
        if arr.ndim != 1 and not arr.is_c_contiguous():
            raise NotImplementedError("put_along_axis(axis=None) needs a contiguous target")
        if indices.ndim != 1:
            raise ValueError("when axis=None, `indices` must have a single dimension.")
        arr, axis = reshape(arr, (arr.size,)), 0
    axis = _norm_axis(axis, arr.ndim)
    off = _along_axis_offsets(arr, indices, axis)
    values = _operand(values)
    if isinstance(values, DeviceArray):
        v = values if values.dtype == arr.dtype else astype(values, arr.dtype)
        v = reshape(copy_(broadcast_to(v, indices.shape)), (-1,))
    else:
        v = DeviceArray.empty((1,), arr.dtype)
        copy_into(v, values)
    check(lib.mdb_scatter_rows(_byref(_rows_view(arr, (), ())), _byref(v.d), _byref(off.d), 0))


# =============================================================================== creation / layout
def _shape_arg(shape):
    return (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)


def full(shape, fill_value, dtype=None, **_kw):
    if dtype is None:
        dtype = np.result_type(fill_value) if not isinstance(fill_value, DeviceArray) else fill_value.dtype
        if isinstance(fill_value, _pyscalar) and not isinstance(fill_value, bool):
            dtype = F64 if isinstance(fill_value, float) else I64
    out = DeviceArray.empty(_shape_arg(shape), dtype)
    copy_into(out, fill_value)
    return out


def zeros(shape, dtype=float, **_kw): return full(shape, 0, dtype=np.dtype(dtype))
def ones(shape, dtype=float, **_kw): return full(shape, 1, dtype=np.dtype(dtype))


def full_like(a, fill_value, dtype=None, **_kw):
    a = asarray(a)
    return full(a.shape, fill_value, dtype=a.dtype if dtype is None else dtype)


def zeros_like(a, dtype=None, **_kw): return full_like(a, 0, dtype)
def ones_like(a, dtype=None, **_kw): return full_like(a, 1, dtype)


def arange(*args, dtype=None, **_kw):
    """np.arange semantics (length, dtype) decided on the host from the scalars, values written by one
    kernel (C ABI mdb_arange): no host array is built."""
    args = [x.item() if isinstance(x, DeviceArray) else x for x in args]
    if not 1 <= len(args) <= 3:
        raise TypeError("arange() requires 1 to 3 positional arguments")
    start, stop, step = (0, args[0], 1) if len(args) == 1 else (args[0], args[1], args[2] if len(args) == 3 else 1)
    if step == 0:
        raise ZeroDivisionError("division by zero")
    dt = np.dtype(dtype) if dtype is not None else np.result_type(*[type(v) if isinstance(v, _pyscalar) else v for v in (start, stop, step)])
    if dt == np.bool_ or dt.kind not in "iuf":
        return DeviceArray.from_numpy(np.arange(*args, dtype=dtype))       # exotic dtypes: NumPy decides
    n = max(0, int(math.ceil((stop - start) / step)))
    out = DeviceArray.empty((n,), dt)
    integral = dt.kind in "iu" and all(isinstance(v, (int, np.integer)) for v in (start, step))
    if n:
        check(lib.mdb_arange(_byref(out.d), float(start), float(step), int(start) if integral else 0,
                             int(step) if integral else 0, 1 if integral else 0))
    return out


def concatenate(arrays, axis=0, **_kw):
    arrays = [asarray(x) for x in arrays]
    if not arrays:
        raise ValueError("need at least one array to concatenate")
    if axis is None:
        arrays, axis = [ravel(x) for x in arrays], 0
    nd = arrays[0].ndim
    if nd == 0:
        raise ValueError("zero-dimensional arrays cannot be concatenated")
    axis = _norm_axis(axis, nd)
    base = list(arrays[0].shape)
    for x in arrays[1:]:
        if x.ndim != nd or any(x.shape[i] != base[i] for i in range(nd) if i != axis):
            raise ValueError("all the input array dimensions except for the concatenation axis "
                             "must match exactly")
    base[axis] = sum(x.shape[axis] for x in arrays)
    out = DeviceArray.empty(base, np.result_type(*[x.dtype for x in arrays]))
    pos = 0
    for x in arrays:
        key = [slice(None)] * nd
        key[axis] = slice(pos, pos + x.shape[axis])
        copy_into(_basic_view(out, tuple(key)), x)
        pos += x.shape[axis]
    return out


def stack(arrays, axis=0, **_kw):
    arrays = [asarray(x) for x in arrays]
    if not arrays:
        raise ValueError("need at least one array to stack")
    if any(x.shape != arrays[0].shape for x in arrays):
        raise ValueError("all input arrays must have the same shape")
    axis = _norm_axis(axis, arrays[0].ndim + 1)
    return concatenate([expand_dims(x, axis) for x in arrays], axis=axis)


def split(ary, indices_or_sections, axis=0):
    ary = asarray(ary)
    axis = _norm_axis(axis, ary.ndim)
    n = ary.shape[axis]
    if isinstance(indices_or_sections, DeviceArray):
        indices_or_sections = indices_or_sections.numpy().tolist()
    if isinstance(indices_or_sections, (int, np.integer)):
        k = int(indices_or_sections)
        if k <= 0 or n % k:
            raise ValueError("array split does not result in an equal division")
        cuts = [n // k * i for i in range(1, k)]
    else:
        cuts = [int(c) for c in indices_or_sections]
    out, prev = [], 0
    for c in cuts + [n]:
        key = [slice(None)] * ary.ndim
        key[axis] = slice(prev, c)
        out.append(_basic_view(ary, tuple(key)))
        prev = c
    return out


def tile(A, reps):
    A = asarray(A)
    if isinstance(reps, DeviceArray):
        reps = reps.numpy().tolist()
    reps = (int(reps),) if isinstance(reps, (int, np.integer)) else tuple(int(r) for r in reps)
    d = max(len(reps), A.ndim)
    reps = (1,) * (d - len(reps)) + reps
    shape = (1,) * (d - A.ndim) + A.shape
    src = reshape(A, shape)
    oshape, ovshape, sshape, sstr = [], [], [], []
    for e, r, s in zip(shape, reps, src.estrides):
        oshape.append(e * r)
        if r == 1:
            ovshape.append(e); sshape.append(e); sstr.append(s)
        elif e == 1:
            ovshape.append(r); sshape.append(r); sstr.append(0)
        else:
            ovshape += [r, e]; sshape += [r, e]; sstr += [0, s]
    if len(ovshape) > _lib.MAX_DIMS:
        raise ValueError("tile: too many dimensions for the device kernel")
    out = DeviceArray.empty(oshape, A.dtype)
    copy_into(out.view(ovshape, c_strides(ovshape)), src.view(sshape, sstr))
    return out


def repeat(a, repeats, axis=None):
    a = asarray(a)
    if axis is None:
        a, axis = ravel(a), 0
    axis = _norm_axis(axis, a.ndim)
    if not isinstance(repeats, (int, np.integer)):
        raise NotImplementedError("repeat with per-element counts is not supported on device")
    r = int(repeats)
    oshape = list(a.shape)
    oshape[axis] *= r
    out = DeviceArray.empty(oshape, a.dtype)
    vshape = list(a.shape[:axis + 1]) + [r] + list(a.shape[axis + 1:])
    if len(vshape) > _lib.MAX_DIMS:
        raise ValueError("repeat: too many dimensions for the device kernel")
    src = a.view(vshape, list(a.estrides[:axis + 1]) + [0] + list(a.estrides[axis + 1:]))
    copy_into(out.view(vshape, c_strides(vshape)), src)
    return out


def vmap(fun):
    """Map `fun` over the leading axis and stack (what NumPy's backend does with
    apply_along_axis, backend/numpy.py:110-122)."""

    def mapped(arr):
        arr = asarray(arr)
        return stack([asarray(fun(arr[i])) for i in range(arr.shape[0])])

    return mapped


# ---- data-dependent helpers, on the device (C ABI mdb_nonzero / mdb_unravel_index / mdb_isin)
def unravel_index(indices, shape):
    idx = asarray(indices)
    if idx.dtype.kind not in "iu":
        raise TypeError("only int indices permitted")
    dims = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s_) for s_ in shape)
    if not 1 <= len(dims) <= _lib.MAX_DIMS:
        raise ValueError(f"unravel_index supports 1..{_lib.MAX_DIMS} dimensions on device")
    flat = idx if idx.is_c_contiguous() else copy_(idx)
    out = DeviceArray.empty((len(dims), flat.size), I64)
    check(lib.mdb_unravel_index(_byref(out.d), _byref(flat.d), len(dims), (C.c_int64 * len(dims))(*dims)))
    return tuple(reshape(getitem(out, d), idx.shape) for d in range(len(dims)))


def argwhere(a):
    a = asarray(a)
    flat = _mask_to_indices(reshape(a, (-1,)) if a.ndim else reshape(a, (1,)))
    if a.ndim == 0:
        return DeviceArray.empty((flat.shape[0], 0), I64)
    if a.ndim == 1:
        return reshape(copy_(flat), (-1, 1))
    coords = DeviceArray.empty((a.ndim, flat.shape[0]), I64)
    if flat.shape[0]:
        check(lib.mdb_unravel_index(_byref(coords.d), _byref(flat.d), a.ndim, (C.c_int64 * a.ndim)(*a.shape)))
    return copy_(coords.T)                      # (n, ndim), C-contiguous like NumPy's


def isin(element, test_elements, assume_unique=False, invert=False, **_kw):
    scalar_in = not isinstance(element, (DeviceArray, np.ndarray, list, tuple))
    e = asarray(element)
    t = asarray(test_elements)
    e_c = e if e.is_c_contiguous() else copy_(e)
    t_c = reshape(t if t.is_c_contiguous() else copy_(t), (-1,))
    out = DeviceArray.empty(e.shape, BOOL)
    if out.size:
        check(lib.mdb_isin(_byref(out.d), _byref(e_c.d), _byref(t_c.d), 1 if invert else 0))
    return out.numpy().reshape(()).item() if scalar_in and not isinstance(element, DeviceArray) else out


def save(file, arr, **kw):
    np.save(file, asarray(arr).numpy(), **kw)


def load(file, **kw):
    return asarray(np.load(file, **kw))


# ---- random: counter-based Philox on the device; the stream cannot match NumPy's MT19937
# (SURVEY 8f rank 3), only the distributions do.
# The stream POSITION lives in device memory (C ABI: offset == MDB_RNG_DEVICE_OFFSET) and is advanced by
# a one-thread kernel behind every draw, so a captured CUDA graph draws fresh numbers on each replay.
_rng_state = {"seed": 0x5EED5EED}
_DEVICE_OFFSET = 2**64 - 1


def seed(s: int):
    _rng_state["seed"] = int(s) & (2**64 - 1)
    _lib.ensure_device()
    check(lib.mdb_random_reset(0))


def _random(shape, normal, dtype=F64):
    out = DeviceArray.empty(shape, dtype)
    check(lib.mdb_random(_byref(out.d), 1 if normal else 0, _rng_state["seed"], _DEVICE_OFFSET))
    return out


def rand(*dims): return _random(tuple(int(d) for d in dims), False)
def randn(*dims): return _random(tuple(int(d) for d in dims), True)


def randint(low, high=None, size=None, dtype=int):
    if high is None:
        low, high = 0, low
    low = low.item() if isinstance(low, DeviceArray) else low
    high = high.item() if isinstance(high, DeviceArray) else high
    if high <= low:
        raise ValueError("low >= high")
    shape = () if size is None else _shape_arg(size)
    out = DeviceArray.empty(shape, np.dtype(dtype))
    if out.dtype.kind not in "iu":
        raise TypeError(f"Unsupported dtype {out.dtype!r} for randint")
    check(lib.mdb_randint(_byref(out.d), int(low), int(high), _rng_state["seed"], _DEVICE_OFFSET))
    return out


def binomial(n, p, size=None):
    """Sum of n Bernoulli(p) draws per output, one launch (C ABI mdb_binomial); p scalar or array."""
    n = int(n.item() if isinstance(n, DeviceArray) else n)
    if n < 0:
        raise ValueError("n < 0")
    shape = () if size is None else _shape_arg(size)
    pd = MdbArray()
    keep = None
    if isinstance(p, (DeviceArray, np.ndarray, list, tuple)):
        p = asarray(p)
        shape = broadcast_shapes([shape, p.shape]) if size is not None else p.shape
        keep = copy_(broadcast_to(p, shape))
        pd = keep.d
    else:
        if not 0.0 <= float(p) <= 1.0:
            raise ValueError("p < 0, p > 1 or p is NaN")
        _fill_imm(pd, float(p))
    out = DeviceArray.empty(shape, I64)
    check(lib.mdb_binomial(_byref(out.d), n, _byref(pd), _rng_state["seed"], _DEVICE_OFFSET))
    return out


def _random_permutation_indices(n: int) -> DeviceArray:
    """uniformly random permutation of 0..n-1 on the device: bitonic sort of (random word, i) keys"""
    bits = DeviceArray.empty((max(n, 1),), np.dtype(np.uint32))
    check(lib.mdb_random_bits(_byref(bits.d), _rng_state["seed"], _DEVICE_OFFSET))
    out = DeviceArray.empty((n,), I64)
    check(lib.mdb_permutation(_byref(out.d), _byref(bits.d)))
    return out


def permutation(x):
    if isinstance(x, (int, np.integer)):
        return _random_permutation_indices(int(x))
    x = asarray(x)
    if x.ndim == 0:
        raise IndexError("x must be an integer or at least 1-dimensional")
    return getitem(x, _random_permutation_indices(x.shape[0]))


def shuffle(x):
    copy_into(x, permutation(x))


def choice(a, size=None, replace=True, p=None):
    pool = arange(int(a)) if isinstance(a, (int, np.integer)) else asarray(a)
    n = pool.shape[0]
    shape = () if size is None else _shape_arg(size)
    k = math.prod(shape)
    if p is None and replace:
        idx = reshape(randint(0, n, size=(k,)), (-1,))
    elif p is None:
        if k > n:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        idx = getitem(_random_permutation_indices(n), slice(0, k))
    else:
        if not replace:
            raise NotImplementedError("weighted sampling without replacement")
        pp = asarray(p)
        if pp.shape != (n,):
            raise ValueError("'a' and 'p' must have same size")
        pp = pp if pp.is_c_contiguous() else copy_(pp)
        cdf = DeviceArray.empty((n,), F64)
        check(lib.mdb_cumsum_f64(_byref(cdf.d), _byref(pp.d)))      # inclusive scan of the weights
        u = _random((k,), False)
        idx = DeviceArray.empty((k,), I64)
        check(lib.mdb_searchsorted_cdf(_byref(idx.d), _byref(cdf.d), _byref(u.d)))
    return reshape(getitem(pool, idx), shape)


# =============================================================================== protocol functions
def tensor_shape(data): return data.shape
def tensor_size(data): return data.size
def tensor_ndim(data): return len(data.shape)
def tensor_dtype(data): return data.dtype
def tensor_item(data): return data.item()
def repr_(data): return data.__repr__()
def len_(data): return data.__len__()


def array_interface(data):
    # raising AttributeError makes NumPy fall through to Tensor.__array__ (reference
    # tensor.py:424-433) instead of treating device memory as host memory
    raise AttributeError("device arrays do not expose __array_interface__")


def array(data, dtype=None, copy=None):
    host = data.numpy()
    if dtype is not None and np.dtype(dtype) != host.dtype:
        if copy is False:
            raise ValueError("attempted cast, but copies are not permitted")
        return host.astype(dtype)
    return host


def as_numpy(a):
    return a.numpy() if isinstance(a, DeviceArray) else np.asarray(a)


def synchronize():
    check(lib.mdb_sync())


# dtype objects are NumPy's (backend/numpy.py:188-202): metadata only
dtype = np.dtype
float64, float32, float16 = np.float64, np.float32, np.float16
uint64, uint32, uint16, uint8 = np.uint64, np.uint32, np.uint16, np.uint8
int64, int32, int16, int8 = np.int64, np.int32, np.int16, np.int8
nan = np.nan
tensor_class = DeviceArray

# the 114 names of the reference's NumPy backend class, in its order.  Names that
# would shadow Python builtins inside this module are defined with a trailing underscore above.
EXPORTED = (
    "tensor_constructor tensor_class "
    "absolute all any argmax argmin argwhere atleast_1d atleast_2d atleast_3d ceil copy cos cosh "
    "exp flatten flip floor invert log logical_not max mean min prod ravel sign sin sinh squeeze "
    "std sum tan tanh transpose add astype broadcast_to dot equal expand_dims floor_divide getitem "
    "greater greater_equal less less_equal logical_and logical_or logical_xor matmul mod multiply "
    "not_equal power reshape subtract tensordot true_divide clip swapaxes where "
    "ones_like ones zeros_like zeros full_like full concatenate index_add isin unravel_index "
    "take_along_axis vmap put_along_axis repeat tile arange stack save load choice rand randint "
    "randn binomial permutation shuffle split "
    "tensor_shape tensor_size tensor_ndim tensor_dtype tensor_item repr len array_interface array "
    "dtype float64 float32 float16 uint64 uint32 uint16 uint8 int64 int32 int16 int8 bool nan "
    "as_numpy"
).split()
assert len(EXPORTED) == 114, len(EXPORTED)
for _n in ("sum", "prod", "max", "min", "any", "all", "len", "repr", "copy"):
    globals()[_n + "_"].__name__ = _n          # op names are taken from __name__ (wrapping.py:150)
_ALIASED = {"bool": np.bool_}
TABLE = {n: _ALIASED[n] if n in _ALIASED else globals().get(n + "_", globals().get(n)) for n in EXPORTED}
assert all(v is not None for v in TABLE.values()), [k for k, v in TABLE.items() if v is None]
