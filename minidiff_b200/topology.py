"""OpNode: one recorded op application, the chain rule, and the reverse sweep.

Same contract as the reference's `minidiff/topology.py` (constructor arguments, `update_grads`,
`toposort`, `backward` with retain_grads / cleanup_mode / allow_higher_order / reset_grads, the
caching hooks) -- but gradient accumulation is device-aware:

  * first-order sweeps (grad mode off, reference topology.py:170) dispatch each input's gradient
    to the op's *fused backward* when it has one: a single launch that evaluates the gradient
    formula, sums away broadcast axes and accumulates into the input's gradient buffer in place
    (replacing grad-lambda -> md.unbroadcast -> out-of-place `grad + new`, topology.py:93-104);
  * a gradient buffer is only written in place while this sweep provably owns it exclusively
    (gradients alias each other in the reference -- `a.grad is b.grad` for `a + b`, read-only
    stride-0 views after `sum` -- SURVEY finding 3), otherwise the reference's out-of-place add runs;
  * higher-order sweeps (allow_higher_order=True) keep building graphs of device ops exactly like
    the reference: no fusion, every gradient op is recorded.
The traversal is iterative, so graph depth is not bounded by Python's recursion limit.
"""
from __future__ import annotations

import minidiff_b200 as md
import minidiff_b200.caching as mdc


class OpNode:
    __slots__ = ("grad_functions", "op_inputs", "op_kwargs", "op_name", "propagate_kwargs",
                 "tensor_inputs", "fused_backward", "_tensor_graph", "_op_ids")

    def __init__(self, forward_func, grad_functions, op_inputs, op_kwargs=None, op_name=None,
                 propagate_kwargs=False, fused_backward=None):
        self.grad_functions = grad_functions
        self.op_inputs = op_inputs
        self.op_kwargs = {} if op_kwargs is None else op_kwargs
        self.op_name = "" if op_name is None else op_name
        self.propagate_kwargs = propagate_kwargs
        self.fused_backward = fused_backward
        Tensor = md.Tensor
        self.tensor_inputs = tensors = [x for x in op_inputs if x.__class__ is Tensor or isinstance(x, Tensor)]
        for t in tensors:
            t.graph_refs += 1
        self._tensor_graph = []
        if not mdc.currently_caching():
            self._op_ids = []
            return
        # structural identity of the sub-graph + nested list of its tensors (topology.py:46-74)
        ids = [-1 if not isinstance(x, md.Tensor) or x.is_leaf else x.op_node._op_ids
               for x in op_inputs]
        ids.append(id(forward_func))
        self._op_ids = tuple(ids)
        seen = set()
        for x in op_inputs:
            if not isinstance(x, md.Tensor) or id(x) in seen:
                continue
            if not x.is_leaf:
                self._tensor_graph.append(x.op_node._tensor_graph)
            self._tensor_graph.append(x)
            seen.add(id(x))

    @property
    def hash(self):
        return hash(self._op_ids)

    # ------------------------------------------------------------------ chain rule
    @staticmethod
    def accumulate(target, contribution, private=False):
        """target.grad (+)= contribution.  `private` says the contribution's buffer was freshly
        produced for this target alone, so later contributions may be added into it in place."""
        if target.grad is None:
            target.grad = contribution
            target._grad_private = contribution if private else None
            return
        if (target._grad_private is target.grad and not md.grad_allowed_()
                and contribution._data.dtype == target.grad._data.dtype):
            target.grad._data += contribution._data        # one in-place kernel, no allocation
            return
        target.grad = target.grad + contribution            # reference path (topology.py:101-104)
        target._grad_private = target.grad if not md.grad_allowed_() else None

    @staticmethod
    def private_grad_buffer(target):
        """The raw gradient buffer of `target` if this sweep may add into it in place, else None."""
        g = target.grad
        if g is not None and target._grad_private is g and not md.grad_allowed_():
            return g._data
        return None

    def update_grads(self, grad):
        fused = self.fused_backward if not md.grad_allowed_() else None
        kwargs = self.op_kwargs if self.propagate_kwargs else {}
        for index, (op_input, grad_function) in enumerate(zip(self.op_inputs, self.grad_functions)):
            if not isinstance(op_input, md.Tensor) or not op_input.allow_grad or grad_function is None:
                continue
            if fused is not None and fused(self, index, op_input, grad):
                continue
            contribution = grad_function(*self.op_inputs, grad, **kwargs)
            if contribution.shape != op_input.shape:
                contribution = md.unbroadcast(contribution, op_input.shape)
            self.accumulate(op_input, contribution)

    # ------------------------------------------------------------------ ordering
    def toposort(self):
        """Tensors below this node in dependency order (inputs before consumers), each once
        (same order as the reference's recursive DFS, topology.py:106-128): iterative post-order
        with one iterator per open node, so graph depth is not bounded by the recursion limit."""
        order, seen = [], set()
        add_seen, emit = seen.add, order.append
        stack = [(None, iter(self.tensor_inputs))]
        while stack:
            tensor, it = stack[-1]
            for child in it:
                cid = id(child)
                if cid in seen:
                    continue
                add_seen(cid)
                node = child.op_node
                if node is None:
                    emit(child)                       # leaf: nothing below it
                else:
                    stack.append((child, iter(node.tensor_inputs)))
                    break
            else:
                stack.pop()
                if tensor is not None:
                    emit(tensor)
        return order

    # ------------------------------------------------------------------ reverse sweep
    def backward(self, seed_grad, retain_grads=False, cleanup_mode="prune",
                 allow_higher_order=False, reset_grads=True):
        if cleanup_mode not in ("keep", "prune", "destroy"):
            raise ValueError(f"Cleanup mode not recognized ({cleanup_mode})")
        if allow_higher_order:  # the graph and intermediate grads are needed again
            retain_grads = True
            if cleanup_mode == "destroy":
                cleanup_mode = "prune"
        if mdc.currently_caching():
            path = []
            for indices in mdc.backward_indices_for_root(self):
                item = self._tensor_graph
                for i in indices:
                    item = item[i]
                path.append(item)
        else:
            path = self.toposort()
        for t in path:
            t._grad_private = None  # in-place accumulation only into buffers born in THIS sweep
            if reset_grads:
                t.grad = None
        # leaves with a grad-ready hook: count how many nodes of this sweep consume them
        waiting = None
        if any(t._grad_hook is not None for t in path):
            waiting = {}
            for node in [self] + [t.op_node for t in path if t.op_node is not None]:
                for x in node.tensor_inputs:
                    if x._grad_hook is not None and x.op_node is None and x.allow_grad:
                        waiting[id(x)] = waiting.get(id(x), 0) + 1

        def consumed(node):
            for x in node.tensor_inputs:
                if id(x) in waiting:
                    waiting[id(x)] -= 1
                    if waiting[id(x)] == 0 and x.grad is not None:
                        x._grad_hook(x)

        with md.enable_grad(allow_higher_order):
            self.update_grads(seed_grad)
            if waiting:
                consumed(self)
            for t in reversed(path):
                node = t.op_node
                if node is None:
                    continue
                node.update_grads(t.grad)
                if waiting:
                    consumed(node)
                if not retain_grads:
                    t.grad = None
                    t._grad_private = None
                if cleanup_mode == "keep":
                    continue
                if cleanup_mode == "destroy":
                    t.wipe()
                    continue
                if t.graph_refs > 0:
                    continue
                for child in node.tensor_inputs:
                    child.graph_refs -= 1
                t.wipe()
        for t in path:
            t._grad_private = None

    def __repr__(self):
        return f"{self.op_name}({', '.join(str(x) for x in self.op_inputs)})"
