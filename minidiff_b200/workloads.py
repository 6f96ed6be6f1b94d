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
    if dp is None:
        with md.no_grad():
            for p in params:
                p -= lr * p.grad
        return loss
    # data parallel: update each parameter as soon as ITS gradient has been averaged, last layer
    # first (the order the backward sweep finished them), while later exchanges are still in flight
    dp.flush()
    with md.no_grad():
        for p in dp.update_order():
            dp.wait(p)
            p -= lr * p.grad
    dp.finish()
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


class HostBatchFeeder:
    """Double-buffered input pipeline for training from host memory: while step i computes on device
    buffers i % 2, batch i + 1 is uploaded from pinned host memory on the library's copy stream
    (mdb_prefetch_h2d).  `next()` returns the (X, Y) Tensors for the step that is about to run."""

    def __init__(self, X_np, Y_np):
        import ctypes as C

        from minidiff_b200.backend._lib import check, lib

        self._C, self._check, self._lib = C, check, lib
        self._host = [self._pin(X_np), self._pin(Y_np)]
        self._dev = [(md.Tensor(md.backend.zeros(X_np.shape, dtype=np.float32)),
                      md.Tensor(md.backend.zeros(Y_np.shape, dtype=np.float32))) for _ in range(2)]
        self._i = 0
        self._issued = False
        self.bytes_per_step = X_np.nbytes + Y_np.nbytes

    def _pin(self, arr):
        C = self._C
        p = C.c_void_p()
        self._check(self._lib.mdb_host_alloc(arr.nbytes, C.byref(p)))
        view = np.frombuffer((C.c_char * arr.nbytes).from_address(p.value), dtype=arr.dtype).reshape(arr.shape)
        view[...] = arr
        return view, p

    def _upload(self, slot):
        for t, (view, _) in zip(self._dev[slot], self._host):
            self._check(self._lib.mdb_prefetch_h2d(t._data.ptr, view.ctypes.data, view.nbytes))

    def next(self, prefetch_following=True):
        slot = self._i % 2
        if not self._issued:                       # very first batch: nothing was prefetched yet
            self._upload(slot)
        self._check(self._lib.mdb_prefetch_wait())  # compute waits for this batch's copy
        self._issued = False
        if prefetch_following:                     # start the next batch's upload behind this step
            self._upload(1 - slot)
            self._issued = True
        self._i += 1
        return self._dev[slot]
