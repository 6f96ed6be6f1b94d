"""Every op the engine exports, with its gradient rules.

Mirrors the *contract* of the reference's `minidiff/ops/definitions.py` (same op names, same
forward semantics == one backend call each, same gradient mathematics, same quirks where user code
can observe them) so results match the reference's NumPy backend; line references point at the
rule being mirrored.  The layout is table-driven rather than one `create_*_op_func` stanza per op,
and ops on the hot path additionally carry a fused device backward (ops/fused.py).
"""
from __future__ import annotations

from builtins import min as _py_min
from math import prod as _prod

import minidiff_b200 as md
import minidiff_b200.backend as backend
from minidiff_b200.ops import fused as _fused
from minidiff_b200.ops import wrapping as _w

__all__ = []


def _export(name, fn):
    globals()[name] = fn
    __all__.append(name)
    return fn


def _unary(name, grad=None, **kw):
    return _export(name, _w.create_unary_op_func(
        forward_func=_w.as_minidiff(getattr(backend, name)), grad=grad,
        fused_backward=getattr(_fused, name, None), **kw))


def _binary(name, grad_x=None, grad_y=None, **kw):
    return _export(name, _w.create_binary_op_func(
        forward_func=_w.as_minidiff(getattr(backend, name)), grad_x=grad_x, grad_y=grad_y,
        fused_backward=getattr(_fused, name, None), **kw))


def _ternary(name, grad_x=None, grad_y=None, grad_z=None, **kw):
    return _export(name, _w.create_ternary_op_func(
        forward_func=_w.as_minidiff(getattr(backend, name)), grad_x=grad_x, grad_y=grad_y,
        grad_z=grad_z, fused_backward=getattr(_fused, name, None), **kw))


# ------------------------------------------------------------------------------ gradient rules
def squeeze_grad(a, grad, axis=None, **kwargs):
    """definitions.py:15-25: put the squeezed-out unit axes back."""
    axes = [i for i, d in enumerate(a.shape) if d == 1] if axis is None else axis
    return expand_dims(grad, axes) if axes else grad


def _contraction_layout(x, y, axes):
    if isinstance(axes, int):
        axes = (tuple(range(x.ndim - axes, x.ndim)), tuple(range(axes)))
    free_x = tuple(i for i in range(x.ndim) if i not in axes[0])
    free_y = tuple(i for i in range(y.ndim) if i not in axes[1])
    return axes, free_x, free_y


def tensordot_grad_x(x, y, grad, axes=2):
    """definitions.py:28-61: contract grad's trailing (y-free) axes with y, then permute the
    result (x-free axes first, x-contracted axes last) back into x's axis order."""
    axes, free_x, free_y = _contraction_layout(x, y, axes)
    g_tail = tuple(range(grad.ndim - len(free_y), grad.ndim))
    res = tensordot(grad, y, axes=(g_tail, free_y))
    perm = [0] * x.ndim
    for pos, ax in enumerate(free_x + tuple(axes[0])):
        perm[ax] = pos
    return md.transpose(res, axes=perm)


def tensordot_grad_y(x, y, grad, axes=2):
    """definitions.py:64-95"""
    axes, free_x, free_y = _contraction_layout(x, y, axes)
    res = tensordot(x, grad, axes=(free_x, tuple(range(len(free_x)))))
    perm = [0] * y.ndim
    for pos, ax in enumerate(tuple(axes[1]) + free_y):
        perm[ax] = pos
    return md.transpose(res, axes=perm)


def max_grad(x, grad, axis=None, **kwargs):
    """definitions.py:98-113 (incl. its behaviour for axis=None / falsy axis)"""
    if axis is None:
        return grad[argmax(x, axis=axis, keepdims=True)]
    if not axis:
        return grad
    where_max = argmax(x, axis=axis, keepdims=True)
    out = md.zeros_like(x)
    md.put_along_axis(out, where_max, grad.reshape(where_max.shape), axis=axis)
    return out


def min_grad(x, grad, axis=None, **kwargs):
    """definitions.py:116-127"""
    where_min = argmin(x, axis=axis, keepdims=True)
    out = md.zeros_like(x)
    md.put_along_axis(out, where_min, grad.reshape(where_min.shape), axis=axis)
    return out


def prod_grad(x, grad, axis=None, **kwargs):
    """definitions.py:130-141: d prod / dx_i = prod / x_i (0 where x_i == 0)."""
    if axis == ():
        return grad.reshape(x.shape)
    total = prod(x, axis=axis, keepdims=True)
    return md.where(x == 0, 0, grad.reshape(total.shape) * total / x)


def transpose_grad(x, grad, axes=None):
    """definitions.py:144-152: inverse permutation (axes elements must offer .item(), as there)."""
    if axes is None:
        return transpose(grad)
    inverse = [-1] * len(axes)
    for i, dim in enumerate(axes):
        inverse[dim.item()] = i
    return transpose(grad, axes=inverse)


def unbroadcast_forward(x, target_shape):
    """definitions.py:157-183: sum away prepended axes, sum-keepdims stretched axes, reshape;
    or broadcast *up* when x is the smaller one."""
    if x.shape == target_shape:
        return x
    lead = tuple(range(x.ndim - len(target_shape)))
    if lead:
        x = x.sum(axis=lead)
    n = _py_min(len(target_shape), x.ndim)
    stretched = tuple(i for i in range(n) if x.shape[i] > 1 and target_shape[i] == 1)
    if stretched:
        x = x.sum(axis=stretched, keepdims=True)
    if x.size == _prod(target_shape):
        return x.reshape(target_shape)
    return broadcast_to(x, target_shape)


def getitem_grad(x, key, grad):
    """definitions.py:186-189: scatter-add the gradient back (duplicates accumulate)."""
    out = md.zeros_like(x)
    md.index_add(out, key, grad)
    return out


def sum_grad(x, grad, axis=None, **kwargs):
    """definitions.py:224-262: re-expand the summed axes by tiling, then permute them home."""
    if isinstance(axis, int):
        axis = tuple(axis)  # TypeError, as in the reference
    if axis is None or not axis:
        return grad
    nd = x.ndim
    hit = [i for i in range(nd) if i in axis]
    k = len(hit)
    tiled = md.tile(grad, [x.shape[i] for i in hit] + [1] * (nd - k))
    perm, moved = [0] * nd, 0
    for i in reversed(range(nd)):
        if moved != k and i == hit[-(moved + 1)]:
            perm[i] = k - 1 - moved
            moved += 1
        else:
            perm[i] = i + moved
    return md.transpose(tiled, axes=perm)


def mean_grad(x, grad, axis=None, **kwargs):
    """definitions.py:192-206 (axis=0 returns the unscaled grad there too)"""
    if axis is None:
        return grad / x.size
    if not axis:
        return grad
    if isinstance(axis, int):
        return grad / x.shape[axis]
    dims = md.Tensor([x.shape[d] for d in axis])
    return sum_grad(x, grad, axis=axis) / prod(dims)


def std_grad(x, grad, axis=None, **kwargs):
    """definitions.py:209-221"""
    if axis is None:
        axis = md.arange(x.ndim)
    if not axis:
        return md.zeros_like(x)
    mu = mean(x, axis=axis)
    n = _prod(d for i, d in enumerate(x.shape) if i in axis)
    return grad * (x - mu) / (std(x, axis=axis, **kwargs) * n)


# ------------------------------------------------------------------------------ unary ops
_unary("absolute", grad=lambda x, grad: grad * sign(x))                 # :266-269
_export("abs", absolute)                                                # :270
for _n in ("all", "any", "argmax", "argmin", "argwhere", "ceil", "floor", "invert",
           "logical_not", "sign"):                                       # non-differentiable
    _unary(_n, is_differentiable=False)
for _n in ("atleast_1d", "atleast_2d", "atleast_3d", "copy"):           # identity gradients
    _unary(_n, grad=lambda x, grad: grad)
_unary("cos", grad=lambda x, grad: grad * -sin(x))                      # :311-314
_unary("cosh", grad=lambda x, grad: grad * sinh(x))                     # :315-318
_unary("exp", grad=lambda x, grad: grad * exp(x))                       # :319-322
_unary("flatten", grad=lambda x, grad, order="C": reshape(grad, x.shape, order=order))
_unary("flip", grad=lambda x, grad, **kw: flip(grad, **kw), propagate_kwargs=True)
_unary("log", grad=lambda x, grad: grad / x)                            # :340-343
_unary("max", grad=max_grad, propagate_kwargs=True)
_unary("mean", grad=mean_grad, propagate_kwargs=True)
_unary("min", grad=min_grad, propagate_kwargs=True)
_unary("prod", grad=prod_grad, propagate_kwargs=True)
_unary("ravel", grad=lambda x, grad, order="C": reshape(grad, x.shape, order=order))
_unary("sin", grad=lambda x, grad: grad * cos(x))                       # :376-379
_unary("sinh", grad=lambda x, grad: grad * cosh(x))
_unary("squeeze", grad=squeeze_grad)
_unary("std", grad=std_grad, propagate_kwargs=True)
_unary("sum", grad=sum_grad, propagate_kwargs=True)
_unary("tan", grad=lambda x, grad: grad * (1 / cos(x) ** 2))            # :408-411
_unary("tanh", grad=lambda x, grad: grad * (1 / cosh(x) ** 2))          # :412-415
_unary("transpose", grad=transpose_grad, propagate_kwargs=True)


def sqrt(a, **kwargs):      # :386-387
    return power(a, 0.5, **kwargs)


def square(a, **kwargs):    # :390-391
    return power(a, 2, **kwargs)


__all__ += ["sqrt", "square"]

# ------------------------------------------------------------------------------ binary ops
_binary("add", grad_x=lambda x, y, grad: grad, grad_y=lambda x, y, grad: grad)
_binary("astype", grad_x=lambda x, dtype, grad: grad.astype(x.dtype))
_binary("broadcast_to", grad_x=lambda x, shape, grad: unbroadcast(grad, x.shape))
_binary("dot", grad_x=lambda x, y, grad: grad * y, grad_y=lambda x, y, grad: grad * x)
_binary("expand_dims", grad_x=lambda x, axis, grad: squeeze(grad, axis=axis))
for _n in ("equal", "floor_divide", "greater", "greater_equal", "less", "less_equal",
           "logical_and", "logical_or", "logical_xor", "not_equal"):
    _binary(_n, is_differentiable=False)
_binary("getitem", grad_x=getitem_grad, op_name="index")
_binary("matmul", grad_x=lambda x, y, grad: matmul(grad, y.T),
        grad_y=lambda x, y, grad: matmul(x.T, grad), tensor_only=True)      # :487-492
_binary("mod", grad_x=lambda x, y, grad: md.where(x % y == 0, 0, grad),
        grad_y=lambda x, y, grad: md.where(x % y == 0, 0, grad))
_binary("multiply", grad_x=lambda x, y, grad: grad * y, grad_y=lambda x, y, grad: grad * x)
_binary("power", grad_x=lambda x, y, grad: grad * y * (x ** (y - 1)),
        grad_y=lambda x, y, grad: grad * log(x) * x**y)                     # :507-511
_binary("reshape", grad_x=lambda x, y, grad: grad.reshape(x.shape))
_binary("subtract", grad_x=lambda x, y, grad: grad, grad_y=lambda x, y, grad: -grad)
_binary("tensordot", grad_x=tensordot_grad_x, grad_y=tensordot_grad_y, tensor_only=True,
        propagate_kwargs=True)
_binary("true_divide", grad_x=lambda x, y, grad: grad / y,
        grad_y=lambda x, y, grad: grad * (-x / y**2))                       # :528-532
_export("unbroadcast", _w.create_binary_op_func(
    forward_func=unbroadcast_forward,
    grad_x=lambda x, shape, grad: broadcast_to(grad, x.shape)))             # :533-536

# ------------------------------------------------------------------------------ ternary ops
_ternary("clip", grad_x=lambda x, a_min, a_max, grad: grad * logical_and(
    1 if a_min is None else x > a_min, 1 if a_max is None else x < a_max))  # :538-547
_ternary("swapaxes", grad_x=lambda x, axis1, axis2, grad, **kw: swapaxes(grad, axis1, axis2, **kw),
         propagate_kwargs=True)
_ternary("where", grad_y=lambda condition, y, z, grad: grad * condition,
         grad_z=lambda condition, y, z, grad: grad * (1 - condition))       # :555-559
del _n
