"""Generate tests/golden/*.npz by running the REAL reference (NumPy backend).

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/{c1..c5,ops}.npz + traces.json

The reference is imported read-only from /root/reference with two shims: a stub
`graphviz` module and a trimmed sys.argv (the reference parses argv at import,
minidiff/backend/__init__.py:13-19).  A logging plugin (oracle/_trace_backend.py) goes through the
reference's own `--backend` loader so the backend-call sequence of each config is recorded as well.
Inputs are produced by the oracle's seeded generators so the oracle, the product tests and these
fixtures all see identical bits.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MINIDIFF_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(HERE, "_stubs"), ROOT, HERE]
sys.argv = [sys.argv[0], "--backend", "oracle._trace_backend"]

import numpy as np  # noqa: E402

import minidiff as md  # noqa: E402  (the real reference)
import minidiff.backend as mdb  # noqa: E402
from oracle import _trace_backend as tr  # noqa: E402
import np_minidiff as orc  # noqa: E402  (input generators only)

assert mdb.multiply.__name__ == "multiply" and mdb.tensor_class is np.ndarray
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
traces = {}


def traced(name, fn):
    tr.CALLS.clear()
    r = fn()
    traces[name] = list(tr.CALLS)
    return r


def relu(h):
    return md.where(h > 0, h, 0)


def mlp(X, ps):
    h = X
    n = len(ps) // 2
    for l in range(n):
        h = h @ ps[2 * l] + ps[2 * l + 1]
        if l < n - 1:
            h = relu(h)
    return h


# ---- C1: README example, fp32
def c1():
    x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
    y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)
    f = traced("c1_fwd", lambda: 2 * y * md.sin(x) - x**2)
    traced("c1_bwd1", lambda: f.backward(allow_higher_order=True))
    r = dict(f=f._data.copy(), dx=x.grad._data.copy(), dy=y.grad._data.copy())
    traced("c1_bwd2", lambda: x.grad.backward())
    r.update(dxx=x.grad._data.copy(), dxy=y.grad._data.copy())
    return r


np.savez(os.path.join(OUT, "c1.npz"), **c1())


# ---- C2: broadcast chain (small and ragged sizes)
def c2(n, m):
    a_np, c_np = orc.config2_inputs(n, m)
    a = md.Tensor(a_np.copy(), allow_grad=True)
    c = md.Tensor(c_np.copy(), allow_grad=True)
    loss = traced("c2_fwd", lambda: md.sum(md.sin(a * c + a) ** 2))
    traced("c2_bwd", lambda: loss.backward())
    return dict(a=a_np, c=c_np, loss=np.asarray(loss._data), da=a.grad._data, dc=c.grad._data)


for n, m in ((64, 48), (257, 131), (1024, 1024)):
    np.savez(os.path.join(OUT, f"c2_{n}x{m}.npz"), **c2(n, m))


# ---- C3: matmul fwd + both gradient GEMMs
def c3(M, K, N):
    A_np = np.random.default_rng(1234).standard_normal((M, K)).astype(np.float32)
    B_np = np.random.default_rng(1235).standard_normal((K, N)).astype(np.float32)
    A = md.Tensor(A_np, allow_grad=True)
    B = md.Tensor(B_np, allow_grad=True)
    C = traced("c3_fwd", lambda: A @ B)
    traced("c3_bwd", lambda: C.backward())
    return dict(A=A_np, B=B_np, C=C._data, dA=A.grad._data, dB=B.grad._data)


for M, K, N in ((48, 40, 56), (256, 384, 128)):
    np.savez(os.path.join(OUT, f"c3_{M}x{K}x{N}.npz"), **c3(M, K, N))

# ---- C4: MLP training step
DIMS = (16, 32, 32, 8)
BATCH = 64


def c4():
    X_np, Y_np = orc.mlp_data(BATCH, DIMS[0], DIMS[-1])
    ps_np = orc.mlp_params(DIMS)
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    loss = traced("c4_fwd", lambda: md.mean((mlp(X, ps) - Y) ** 2))
    traced("c4_bwd", lambda: loss.backward())
    grads = [p.grad._data.copy() for p in ps]

    def upd():
        with md.no_grad():
            for p in ps:
                p -= 0.01 * p.grad

    traced("c4_update", upd)
    r = dict(loss=np.asarray(loss._data))
    for i, (g, p) in enumerate(zip(grads, ps)):
        r[f"g{i}"] = g
        r[f"p{i}"] = p._data
    return r


np.savez(os.path.join(OUT, "c4.npz"), **c4())


# ---- C5: Hessian-vector product (unreduced loss; SURVEY finding 1)
def c5():
    X_np, Y_np = orc.mlp_data(BATCH, DIMS[0], DIMS[-1])
    ps_np = orc.mlp_params(DIMS)
    vs_np = [np.random.default_rng(100 + i).standard_normal(p.shape).astype(np.float32)
             for i, p in enumerate(ps_np)]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    vs = [md.Tensor(v) for v in vs_np]
    out = mlp(X, ps)
    L = ((out - Y) ** 2) / float(out.size)
    traced("c5_bwd1", lambda: L.backward(allow_higher_order=True))
    g1 = [p.grad._data.copy() for p in ps]

    def dot():
        s = None
        for p, v in zip(ps, vs):
            t = md.sum(p.grad * v)
            s = t if s is None else s + t
        return s

    s = traced("c5_dot", dot)
    traced("c5_bwd2", lambda: s.backward())
    r = {}
    for i, (g, p, v) in enumerate(zip(g1, ps, vs_np)):
        r[f"g{i}"] = g
        r[f"hv{i}"] = p.grad._data.copy()
        r[f"v{i}"] = v
    return r


np.savez(os.path.join(OUT, "c5.npz"), **c5())

# ---- per-op vectors: forward + gradients of sum(op(...) * w) for seeded fp32 inputs
rng = np.random.default_rng(7)


def f32(*shape, lo=None, hi=None):
    v = rng.standard_normal(shape).astype(np.float32)
    if lo is not None:
        v = (np.abs(v) + lo).astype(np.float32)
    return v


OPS = {}


def case(name, fn, *inputs):
    ts = [md.Tensor(v.copy(), allow_grad=True) if isinstance(v, np.ndarray) and v.dtype == np.float32
          else v for v in inputs]
    out = fn(*ts)
    rec = {"out": np.asarray(out._data)}
    if out.allow_grad and not out.is_leaf:
        w = md.Tensor(np.random.default_rng(11).standard_normal(out.shape).astype(out._data.dtype))
        try:
            md.sum(out * w).backward()
        except Exception as e:  # reference bug on this path: pin the error type instead
            rec["raises"] = np.array(type(e).__name__)
        else:
            for i, t in enumerate(ts):
                if isinstance(t, md.Tensor) and t.grad is not None:
                    rec[f"g{i}"] = np.asarray(t.grad._data)
    for i, v in enumerate(inputs):
        if isinstance(v, np.ndarray):
            rec[f"in{i}"] = v
    for k, v in rec.items():
        OPS[f"{name}/{k}"] = v


A, B = f32(5, 7), f32(5, 7)
col, row = f32(5, 1), f32(1, 7)
vec = f32(7)
pos = f32(5, 7, lo=0.5)
for nm in ("add", "subtract", "multiply", "true_divide"):
    case(nm, getattr(md, nm), A, B)
    case(nm + "_bcast_col_row", getattr(md, nm), col, row)
    case(nm + "_bcast_vec", getattr(md, nm), A, vec)
    case(nm + "_scalar_l", lambda t, f=getattr(md, nm): f(2.5, t), A)
    case(nm + "_scalar_r", lambda t, f=getattr(md, nm): f(t, 2.5), A)
for e in (2, 1, 0, 0.5, -1, 3, 2.5):
    case(f"power_{e}", lambda t, e=e: t**e, pos)
case("power_tt", md.power, pos, f32(5, 7))
for nm in ("sin", "cos", "tan", "exp", "sinh", "cosh", "tanh", "absolute", "copy"):
    case(nm, getattr(md, nm), A)
case("log", md.log, pos)
case("sqrt", md.sqrt, pos)
case("square", md.square, A)
case("neg", lambda t: -t, A)
case("where_relu", lambda t: md.where(t > 0, t, 0), A)
case("where_tt", lambda t, u: md.where(t > u, t, u), A, B)
case("clip", lambda t: md.clip(t, -0.5, 0.5), A)
case("clip_lo", lambda t: md.clip(t, 0, None), A)
case("mask_mul", lambda t: t * (t > 0), A)
for nm in ("greater", "greater_equal", "less", "less_equal", "equal", "not_equal"):
    case(nm, getattr(md, nm), A, B)
case("sum_all", md.sum, A)
case("mean_all", md.mean, A)
case("sum_ax0", lambda t: md.sum(t, axis=(0,)), A)
case("sum_ax1_keep", lambda t: md.sum(t, axis=(1,), keepdims=True), A)
case("sum_ax02", lambda t: md.sum(t, axis=(0, 2)), f32(3, 4, 5))
case("max_all", md.max, A)
case("max_ax1", lambda t: md.max(t, axis=1), A)
case("min_ax1", lambda t: md.min(t, axis=1), A)
case("prod_ax0", lambda t: md.prod(t, axis=0), A)
case("transpose", md.transpose, A)
case("T_matmul", lambda t, u: t.T @ u, A, f32(5, 3))
case("matmul", md.matmul, A, f32(7, 3))
case("reshape", lambda t: md.reshape(t, (7, 5)), A)
case("broadcast_to", lambda t: md.broadcast_to(t, (4, 5, 7)), A)
case("expand_dims", lambda t: md.expand_dims(t, 1), A)
case("squeeze", md.squeeze, f32(5, 1, 7))
case("swapaxes", lambda t: md.swapaxes(t, 0, 2), f32(3, 4, 5))
case("flip", lambda t: md.flip(t, axis=1), A)
case("ravel", md.ravel, A)
case("flatten", md.flatten, A)
case("getitem_slice", lambda t: t[1:4, ::2], A)
case("getitem_int", lambda t: t[2], A)
case("tensordot", lambda t, u: md.tensordot(t, u, axes=1), A, f32(7, 3))
case("dot", md.dot, vec, f32(7))
case("mod", lambda t: md.mod(t, 0.75), A)
case("floor", md.floor, A)
case("ceil", md.ceil, A)
case("sign", md.sign, A)
case("std_ax1", lambda t: md.std(t, axis=(1,)), A)
case("chain_bcast", lambda a, c: md.sin(a * c + a) ** 2, col, row)
np.savez(os.path.join(OUT, "ops.npz"), **OPS)

with open(os.path.join(OUT, "traces.json"), "w") as fh:
    json.dump(traces, fh, indent=0)
print("golden written:", sorted(os.listdir(OUT)))
print({k: len(v) for k, v in traces.items()})
