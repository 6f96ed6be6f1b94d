"""Fused first-order backward handlers: "binding op funcs to device kernels".

Each handler has the signature `(node, index, op_input, grad) -> bool` and, when it returns True,
has already done for input `index` what the reference does with a chain of backend calls:

    grad_function(...)            # 1-3 elementwise launches   (ops/definitions.py grad lambdas)
    md.unbroadcast(g, shape)      # 1-2 reductions + reshape   (ops/definitions.py:157-183)
    op_input.grad = op_input.grad + g   # 1 more launch + a fresh buffer (topology.py:101-104)

as ONE launch of `mdb_elementwise_reduce` / `mdb_elementwise` / `mdb_gemm` that writes (or adds
into) the input's gradient buffer.  The fused kernels round every intermediate exactly like the
unfused chain (csrc/ew_ops.cuh), so results are bit-identical to the reference call sequence for
elementwise work and differ only by summation order for the reductions.

Handlers return False whenever the fast form does not apply (non-fp32 storage, up-broadcasts,
tensor exponents, ...) and the engine falls back to the reference call chain.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import minidiff_b200 as md
from minidiff_b200.backend import _lib
from minidiff_b200.backend import functions as F
from minidiff_b200.backend._lib import OP, MdbArray, check, lib
from minidiff_b200.backend.device_array import BOOL, F32, DeviceArray
from minidiff_b200.topology import OpNode

_scalar = (int, float)


def _raw(x):
    return x._data if isinstance(x, md.Tensor) else x


def _ok_operand(x) -> bool:
    if isinstance(x, DeviceArray):
        return x.dtype == F32 or x.dtype == BOOL
    return isinstance(x, _scalar) and not isinstance(x, bool)


def _reduces_to(shape, tshape) -> bool:
    """True if a value of `shape` un-broadcasts to `tshape` purely by summing axes."""
    if len(tshape) > len(shape):
        return False
    lead = len(shape) - len(tshape)
    return all(t == s or t == 1 for s, t in zip(shape[lead:], tshape))


def contribute(target, op: str, *operands) -> bool:
    """target.grad (+)= unbroadcast(op(*operands), target.shape), one launch."""
    tdata = target._data
    if tdata.dtype != F32:
        return False
    shape, any_f32, same = None, False, True
    for o in operands:
        c = o.__class__
        if c is DeviceArray:
            dt = o.dtype
            if dt == F32:
                any_f32 = True
            elif dt != BOOL:
                return False
            if shape is None:
                shape = o.shape
            elif o.shape != shape:
                same = False
        elif not ((c is int or c is float)):
            if not _ok_operand(o):
                return False
    if not any_f32:
        return False
    if not same:
        shape = F.broadcast_shapes([o.shape for o in operands if o.__class__ is DeviceArray])
    tshape = tdata.shape
    dst = OpNode.private_grad_buffer(target)
    if shape == tshape:
        if dst is not None and op == "MUL":
            if not F._ew_into("FMA", dst, (dst, *operands)):          # dst = dst + a*b, in place
                F._launch_ew("FMA", dst, [dst, *operands])
            return True
        fresh = F._ew(op, F32, *operands)
        OpNode.accumulate(target, md.Tensor._wrap(fresh), private=True)
        return True
    if not _reduces_to(shape, tshape):
        return False
    out = dst if dst is not None else DeviceArray.empty(tshape, F32)
    n = len(operands)
    descs = (MdbArray * n)()
    for i, o in enumerate(operands):
        if isinstance(o, DeviceArray):
            descs[i] = o.d
        else:
            F._fill_imm(descs[i], o)
    check(lib.mdb_elementwise_reduce(OP[op], C.byref(out.d), n, descs, 1 if dst is not None else 0))
    if dst is None:
        OpNode.accumulate(target, md.Tensor._wrap(out), private=True)
    return True


# ---------------------------------------------------------------------------- per-op handlers
def multiply(node, index, op_input, grad):
    other = _raw(node.op_inputs[1 - index])
    return contribute(op_input, "MUL", grad._data, other)


def add(node, index, op_input, grad):
    if grad.shape == op_input.shape:
        return False          # identity gradient: alias it like the reference does (no launch)
    return contribute(op_input, "COPY", grad._data)


def subtract(node, index, op_input, grad):
    if index == 0:
        return add(node, index, op_input, grad)
    return contribute(op_input, "NEG", grad._data)


def true_divide(node, index, op_input, grad):
    x, y = (_raw(v) for v in node.op_inputs)
    if index == 0:
        return contribute(op_input, "DIV", grad._data, y)
    return contribute(op_input, "DIV_BWD_Y", grad._data, x, y)


def power(node, index, op_input, grad):
    x, y = node.op_inputs
    if index != 0 or not isinstance(y, _scalar) or isinstance(y, bool):
        return False
    return contribute(op_input, "POW_BWD", grad._data, x._data, y)


def _unary(op):
    def handler(node, index, op_input, grad):
        return contribute(op_input, op, grad._data, op_input._data)

    return handler


sin, cos, exp, log, tanh = (_unary(n) for n in ("SIN_BWD", "COS_BWD", "EXP_BWD", "LOG_BWD", "TANH_BWD"))


def where(node, index, op_input, grad):
    if index != 1:
        return False          # grad_z promotes to float64 in the reference: keep its chain
    cond = _raw(node.op_inputs[0])
    if not isinstance(cond, DeviceArray) or cond.dtype != BOOL:
        return False
    return contribute(op_input, "MUL", grad._data, cond)


def matmul(node, index, op_input, grad):
    """dX = dC @ Y^T and dY = X^T @ dC straight from transposed *views* (no copies), accumulated
    into the gradient buffer by the GEMM epilogue when the buffer is privately owned."""
    x, y = (v._data for v in node.op_inputs)
    g = grad._data
    if x.ndim != 2 or y.ndim != 2 or g.ndim != 2 or not (x.dtype == y.dtype == g.dtype == F32):
        return False
    a, b = (g, y.T) if index == 0 else (x.T, g)
    if (a.shape[0], b.shape[1]) != op_input._data.shape:
        return False
    dst = OpNode.private_grad_buffer(op_input)
    if dst is not None and dst.dtype == F32 and dst.is_c_contiguous():
        F._gemm(a, b, out=dst, accumulate=True)
        return True
    OpNode.accumulate(op_input, md.Tensor(F._gemm(a, b)), private=True)
    return True
