"""The five BASELINE configurations written against the engine's public API (what a minidiff user
would write), plus their synthetic-input generators.  Used by bench.py and the examples; NumPy is
used for input generation only."""
from __future__ import annotations

import numpy as np

import minidiff_b200 as md

E_BYTES = (1 << 26) * 4                       # one 2^26-element fp32 tensor ("E")
MLP_DIMS = (1024, 4096, 4096, 1024)


# ------------------------------------------------------------------ inputs
def c2_inputs(n=8192, m=8192, seed=1234):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, 1)).astype(np.float32),
            rng.standard_normal((1, m)).astype(np.float32))


def c3_inputs(n=8192):
    return (np.random.default_rng(1234).standard_normal((n, n), dtype=np.float32),
            np.random.default_rng(1235).standard_normal((n, n), dtype=np.float32))


def mlp_params(dims=MLP_DIMS, seed0=2):
    ps, s = [], seed0
    for fi, fo in zip(dims[:-1], dims[1:]):
        ps.append((np.random.default_rng(s).standard_normal((fi, fo)) / np.sqrt(fi)).astype(np.float32))
        ps.append((np.random.default_rng(s + 1).standard_normal((fo,)) / np.sqrt(fo)).astype(np.float32))
        s += 2
    return ps


def mlp_data(batch, d_in, d_out, seed=0):
    X = np.random.default_rng(seed).standard_normal((batch, d_in), dtype=np.float32)
    Y = np.random.default_rng(seed + 1).standard_normal((batch, d_out), dtype=np.float32)
    return X, Y


# ------------------------------------------------------------------ C2
def c2_step(a, c):
    """loss = sum(sin(a*c + a)**2); backward with un-broadcast gradient sums."""
    loss = md.sum(md.sin(a * c + a) ** 2)
    loss.backward()
    return loss


C2_ALGORITHMIC_BYTES = 26 * E_BYTES            # fwd 8E + bwd 18E of the reference chain


# ------------------------------------------------------------------ C3
def c3_step(A, B):
    C = A @ B
    C.backward()
    return C


def c3_flops(n=8192):
    return 3 * 2.0 * n ** 3


# ------------------------------------------------------------------ C4
def relu(h):
    return md.where(h > 0, h, 0)           # the reference has no relu op (SURVEY finding 2)


def mlp_forward(X, params):
    h = X
    n = len(params) // 2
    for l in range(n):
        h = h @ params[2 * l] + params[2 * l + 1]
        if l < n - 1:
            h = relu(h)
    return h


def mlp_train_step(X, Y, params, lr=0.01, dp=None):
    """Full training step: forward, mean-MSE, backward, (DP: all-reduce grads), SGD under no_grad."""
    loss = md.mean((mlp_forward(X, params) - Y) ** 2)
    loss.backward()
    if dp is not None:
        dp.finish()
    with md.no_grad():
        for p in params:
            p -= lr * p.grad
    return loss


def mlp_flops_per_sample(dims=MLP_DIMS):
    """fwd + dW for every layer, dX for all but the first (X has allow_grad=False)."""
    layers = list(zip(dims[:-1], dims[1:]))
    fwd = sum(i * o for i, o in layers)
    dx = sum(i * o for i, o in layers[1:])
    return 2.0 * (2 * fwd + dx)


# ------------------------------------------------------------------ C5
def hvp(X, Y, params, vs):
    """Hessian-vector product through an UNREDUCED loss (SURVEY finding 1)."""
    out = mlp_forward(X, params)
    L = ((out - Y) ** 2) / float(out.size)
    L.backward(allow_higher_order=True)
    s = None
    for p, v in zip(params, vs):
        t = md.sum(p.grad * v)
        s = t if s is None else s + t
    s.backward()
    return [p.grad for p in params]
