"""Gradient checking utilities (reference minidiff/utils.py:104-197).

`calculate_finite_differences` / `compute_grads` keep the reference's signatures and semantics
(central differences of a scalar-valued `func`, one perturbed copy of the input per element, all
copies evaluated through `md.vmap`, Tensors in `exclude` skipped) and run entirely on the device
backend -- they are how first- and second-order gradients of the CUDA kernels are verified
(tests/), exercising tile / fancy setitem / vmap / reshape on DeviceArrays.
The reference's graphviz drawing helper (utils.py:17-101) is visualisation and out of scope.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

import minidiff_b200 as md


def calculate_finite_differences(*input_tensors, func, h=1e-7, exclude=None):
    excluded = {id(x) for x in (exclude or [])}
    results = []
    with md.no_grad():
        for pos, t in enumerate(input_tensors):
            if not isinstance(t, md.Tensor) or not t.allow_grad or id(t) in excluded:
                results.append(None)
                continue
            n = t.size
            before, after = input_tensors[:pos], input_tensors[pos + 1:]

            def at(shifted, _b=before, _a=after):
                return func(*_b, shifted, *_a)

            probe = md.vmap(at)
            # row k of `plus` / `minus` is a copy of t with element k nudged by +-h
            coords = md.Tensor(np.array(list(np.ndindex(t.shape)), dtype=np.int64).reshape(n, t.ndim))
            where = (md.arange(n), *[coords[:, d] for d in range(t.ndim)])
            reps = (n,) + (1,) * t.ndim
            plus = md.tile(t.detach().copy(), reps)
            minus = md.tile(t.detach().copy(), reps)
            plus[*where] += h
            minus[*where] -= h
            slope = (probe(plus) - probe(minus)) / (2 * h)
            results.append(slope.reshape(t.shape))
    return results


def compute_grads(*input_tensors, func, h=1e-7, exclude=None):
    """(finite-difference grads, autodiff grads) for fresh copies of the inputs (utils.py:163-197)."""
    excluded = {id(x) for x in (exclude or [])}
    copies, copied_exclude = [], []
    for t in input_tensors:
        c = t.copy().detach(allow_grad=True) if isinstance(t, md.Tensor) else deepcopy(t)
        copies.append(c)
        if id(t) in excluded:
            copied_exclude.append(c)
    func(*copies).backward(retain_grads=True)
    automatic = [t.grad if isinstance(t, md.Tensor) else None for t in copies]
    manual = calculate_finite_differences(*copies, func=func, h=h, exclude=copied_exclude)
    return manual, automatic
