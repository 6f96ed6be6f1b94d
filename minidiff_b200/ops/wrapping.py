"""Turning backend functions into differentiable Tensor ops.

Same helpers, names and argument meaning as the reference's `minidiff/ops/wrapping.py`
(`as_minidiff`, `create_op_func`, `create_stateful_op_func`, the unary/binary/ternary sugar, the
decorator forms, `OpClass` & friends), so custom ops written against the reference keep working.
One addition binds an op to device kernels: `fused_backward=` names a handler that computes an
input's gradient, un-broadcasts it and accumulates it in place with a single launch; the engine
uses it on first-order sweeps (see topology.OpNode.update_grads).
"""
from __future__ import annotations

import minidiff_b200 as md
from minidiff_b200.topology import OpNode

__all__ = [
    "OpClass", "UnaryOpClass", "BinaryOpClass", "TernaryOpClass", "op_func", "unary_op_func",
    "binary_op_func", "ternary_op_func", "as_minidiff", "create_op_func", "create_stateful_op_func",
    "create_unary_op_func", "create_binary_op_func", "create_ternary_op_func",
]


def _tracks_grad(op_inputs) -> bool:
    """Output tracks gradients iff recording is on and some Tensor input does (wrapping.py:17-25)."""
    if not md.grad_allowed_():
        return False
    for x in op_inputs:
        if isinstance(x, md.Tensor) and x.allow_grad:
            return True
    return False


def _check_inputs(op_inputs, tensor_only: bool) -> None:
    """wrapping.py:28-44: all inputs must be Tensors (tensor_only) or at least ... the first
    deciding element settles it, exactly like the reference's early-exit loop."""
    ok = False
    for x in op_inputs:
        ok = isinstance(x, md.Tensor)
        if ok != tensor_only:
            break
    if ok:
        return
    raise ValueError("This function only supports minidiff Tensors" if tensor_only else
                     "This function requires at least one minidiff Tensor argument")


class OpClass:
    """Stateful op protocol (wrapping.py:47-76): `create_forward()` / `create_grads()`."""

    def create_forward(self):
        raise NotImplementedError

    def create_grads(self):
        raise NotImplementedError


class UnaryOpClass(OpClass):
    pass


class BinaryOpClass(OpClass):
    pass


class TernaryOpClass(OpClass):
    pass


_CONTAINERS = (tuple, list, dict)


def _unwrap_args(args):
    """try_unwrap over positional arguments, without the generic recursion for the common case
    (Tensors and plain scalars)."""
    Tensor, unwrap = md.Tensor, md.try_unwrap
    return [a._data if a.__class__ is Tensor else (unwrap(a) if isinstance(a, _CONTAINERS) or isinstance(a, Tensor) else a)
            for a in args]


def as_minidiff(func):
    """Raw-array function -> Tensor function (wrapping.py:117-134): unwrap Tensors in args and
    kwargs, make ONE backend call (== one kernel launch on the compute stream), wrap the result."""

    def wrapper(*args, **kwargs):
        allow_grad = _tracks_grad(args)
        if kwargs:
            out = func(*_unwrap_args(args), **md.try_unwrap(kwargs))
        else:
            out = func(*_unwrap_args(args))
        return md.Tensor(out, allow_grad=allow_grad)

    wrapper.__name__ = func.__name__
    wrapper.__qualname__ = getattr(func, "__qualname__", func.__name__)
    wrapper._mdb_raw = func          # create_op_func calls the backend function directly (one wrap, not two)
    return wrapper


def _finish(output, allow_grad, node_factory):
    if output.op_node is not None:       # produced by another op: adopt a detached alias
        output = output.detach()
    output.allow_grad = allow_grad
    if node_factory is not None and allow_grad and md.grad_allowed_():
        output.op_node = node_factory()
    return output


def create_op_func(forward_func, grad_funcs, propagate_kwargs=False, is_differentiable=True,
                   tensor_only=False, op_name=None, fused_backward=None):
    """wrapping.py:137-178.  `fused_backward(node, index, op_input, grad) -> bool` is optional."""
    if not is_differentiable:
        grad_funcs = [None] * len(grad_funcs)
    if op_name is None:
        op_name = forward_func.__name__

    raw = getattr(forward_func, "_mdb_raw", None)
    grad_mode = md.grad_allowed_

    def minidiff_func(*op_inputs, **op_kwargs):
        if raw is not None:
            # forward_func is as_minidiff(raw).  Same checks and results as the general path below, done
            # in ONE pass over the inputs: validation (wrapping.py:28-44), the allow_grad decision
            # (wrapping.py:17-25) and the unwrapping, then one backend call and one Tensor.
            Tensor = md.Tensor
            args, ok, tracks = [], False, False
            decided = False
            for a in op_inputs:
                if a.__class__ is Tensor or isinstance(a, Tensor):
                    args.append(a._data)
                    if a._allow_grad:
                        tracks = True
                    is_t = True
                else:
                    args.append(md.try_unwrap(a) if isinstance(a, _CONTAINERS) else a)
                    is_t = False
                if not decided:
                    ok = is_t
                    decided = is_t != tensor_only
            if not ok:
                raise ValueError("This function only supports minidiff Tensors" if tensor_only else
                                 "This function requires at least one minidiff Tensor argument")
            allow_grad = tracks and grad_mode()
            out = raw(*args, **md.try_unwrap(op_kwargs)) if op_kwargs else raw(*args)
            output = Tensor._wrap(out, allow_grad) if out.__class__ is Tensor._raw_class else Tensor(out, allow_grad=allow_grad)
            if is_differentiable and allow_grad:
                output.op_node = OpNode(forward_func, grad_funcs, op_inputs, op_kwargs, op_name,
                                        propagate_kwargs, fused_backward)
            return output
        _check_inputs(op_inputs, tensor_only)
        allow_grad = _tracks_grad(op_inputs)
        output = forward_func(*op_inputs, **op_kwargs)
        factory = None
        if is_differentiable:
            factory = lambda: OpNode(forward_func, grad_funcs, op_inputs, op_kwargs, op_name,  # noqa: E731
                                     propagate_kwargs, fused_backward)
        return _finish(output, allow_grad, factory)

    minidiff_func.__name__ = op_name
    minidiff_func.__qualname__ = f"<op func '{op_name}'>"
    return minidiff_func


def create_stateful_op_func(op_class, propagate_kwargs=False, tensor_only=False, op_name=None):
    """wrapping.py:181-217: a fresh `op_class()` per call supplies forward and grads."""
    if op_name is None:
        op_name = op_class.__name__

    def minidiff_func(*op_inputs, **op_kwargs):
        _check_inputs(op_inputs, tensor_only)
        allow_grad = _tracks_grad(op_inputs)
        instance = op_class()
        forward = instance.create_forward()
        output = forward(*op_inputs, **op_kwargs)
        factory = lambda: OpNode(forward, instance.create_grads(), op_inputs, op_kwargs, op_name,  # noqa: E731
                                 propagate_kwargs, getattr(instance, "fused_backward", None))
        return _finish(output, allow_grad, factory)

    minidiff_func.__name__ = op_name
    minidiff_func.__qualname__ = f"<stateful op func '{op_name}'>"
    return minidiff_func


def create_unary_op_func(forward_func, grad=None, **kwargs):
    return create_op_func(forward_func, [grad], **dict(kwargs, tensor_only=True))


def create_binary_op_func(forward_func, grad_x=None, grad_y=None, **kwargs):
    return create_op_func(forward_func, [grad_x, grad_y], **kwargs)


def create_ternary_op_func(forward_func, grad_x=None, grad_y=None, grad_z=None, **kwargs):
    return create_op_func(forward_func, [grad_x, grad_y, grad_z], **kwargs)


def _decorator(creator):
    def deco(**kwargs):
        return lambda func: creator(forward_func=func, **kwargs)

    return deco


op_func = _decorator(create_op_func)
unary_op_func = _decorator(create_unary_op_func)
binary_op_func = _decorator(create_binary_op_func)
ternary_op_func = _decorator(create_ternary_op_func)
