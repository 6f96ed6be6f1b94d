"""User-level fused device ops on the reference's stateful-op protocol (SURVEY 8f-4:
create_stateful_op_func / OpClass, reference ops/wrapping.py:47-76,181-217): md.relu,
md.linear_relu, md.linear.  Run under BOTH engines (this repo's and the unmodified reference's,
through minidiff_b200.ops.fused_ops.make_ops(md)) against the oracle's 3-op composition
relu(X @ W + b), relu = where(h > 0, h, 0): forward bit-exact, gradients rtol 1e-4."""
import numpy as np
import pytest

import np_minidiff as orc
from test_gpu_engine import _close_rms, close, md, mlp, run_c4  # noqa: F401  (md is the engine fixture)

pytestmark = pytest.mark.gpu


def fused_ops(md):
    if md.__name__ == "minidiff_b200":
        return md.relu, md.linear_relu, md.linear
    from minidiff_b200.ops.fused_ops import make_ops

    return make_ops(md)


def launches():
    from minidiff_b200.backend._lib import lib

    return int(lib.mdb_launch_count())


def fused_mlp(md, X, ps):
    relu, linear_relu, linear = fused_ops(md)
    h = linear_relu(X, ps[0], ps[1])
    h = linear_relu(h, ps[2], ps[3])
    return linear(h, ps[4], ps[5])


@pytest.mark.parametrize("dims,B", [((256, 512, 512, 384), 1024), ((16, 32, 32, 8), 64)])
def test_linear_relu_mlp_matches_the_composition(md, dims, B):
    """3-layer MLP written with linear_relu / linear: forward bit-identical to the where-ReLU
    composition on the same engine, loss and all six gradients within rtol 1e-4 of the oracle."""
    X_np, Y_np = orc.mlp_data(B, dims[0], dims[-1])
    ps_np = orc.mlp_params(dims)
    want = orc.config4_step(X_np, Y_np, ps_np)
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    out = fused_mlp(md, X, ps)
    ref_out = mlp(md, X, [md.Tensor(p.copy()) for p in ps_np])
    assert np.array_equal(out.as_numpy(), ref_out.as_numpy()), "fused forward must be bit-identical"
    loss = md.mean((out - Y) ** 2)
    loss.backward()
    close(loss, want["loss"])
    for i in range(6):
        _close_rms(ps[i].grad.as_numpy(), want["grads"][i], f"grad {i}")
    assert out.op_node is not None and out.op_node.op_name == "linear"


def test_linear_relu_baseline_dims_vs_oracle(md):
    """BASELINE layer dims (1024-4096-4096-1024), 2048 kink-safe rows: every GEMM of the fused step
    runs on the CTA-pair kernel with its epilogue; parity vs the oracle composition."""
    from test_gpu_engine import BASELINE_DIMS, _c4_baseline_case, _gemm_paths

    X_np, Y_np, ps_np, want, truth = _c4_baseline_case()
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    _gemm_paths(reset=True)
    l0 = launches()
    loss = md.mean((fused_mlp(md, X, ps) - Y) ** 2)
    loss.backward()
    n_launch = launches() - l0
    paths = _gemm_paths()
    assert paths["pair"] + paths["pair_streamk"] == 8 and paths["simt"] == 0, paths
    close(loss, want["loss"], rtol=1e-5)
    for i in range(6):
        _close_rms(ps[i].grad.as_numpy(), want["grads"][i], f"grad {i} vs oracle")
        _close_rms(ps[i].grad.as_numpy(), truth["grads"][i], f"grad {i} vs float64")
    if md.__name__ == "minidiff_b200":
        # 3 fused forward GEMMs + loss chain (sub, pow, mean) + backward: seed/mean/pow/sub chain,
        # 5 gradient GEMMs, 3 bias column sums, ONE masking pass (layer 2's; layer 1's mask rides in
        # the epilogue of layer 2's dX GEMM)
        assert n_launch <= 20, n_launch


def test_relu_op_forward_and_backward(md):
    relu, _, _ = fused_ops(md)
    rng = np.random.default_rng(3)
    x_np = rng.standard_normal((300, 257)).astype(np.float32)
    x_np[5, 7] = 0.0
    x_np[6, 8] = np.nan
    x = md.Tensor(x_np, allow_grad=True)
    l0 = launches()
    y = relu(x)
    assert launches() - l0 == 1
    with np.errstate(invalid="ignore"):
        want = np.where(x_np > 0, x_np, 0).astype(np.float32)
    assert np.array_equal(y.as_numpy(), want)
    up = rng.standard_normal(x_np.shape).astype(np.float32)
    (y * md.Tensor(up)).backward()
    with np.errstate(invalid="ignore"):
        assert np.array_equal(x.grad.as_numpy(), up * (x_np > 0))


def test_stateful_op_protocol_fresh_instance_per_call_and_shared_state(md):
    """create_stateful_op_func contract (reference wrapping.py:181-217): a new OpClass instance per
    call; create_forward() then create_grads() once each; state made by the forward is visible to the
    gradient functions; output detached / allow_grad follows the inputs; no node under no_grad."""
    wrapping = md.ops.wrapping
    made = []

    class Scale(wrapping.OpClass):
        def __init__(self):
            made.append(self)
            self.calls = []

        def create_forward(self):
            self.calls.append("forward")

            def forward(x, k=2.0):
                self.k = k
                return md.Tensor(x._data * k)

            return forward

        def create_grads(self):
            self.calls.append("grads")
            return [lambda x, grad, k=2.0: grad * self.k]

    scale = wrapping.create_stateful_op_func(Scale, propagate_kwargs=True, tensor_only=True)
    x = md.Tensor(np.arange(6, dtype=np.float32).reshape(2, 3), allow_grad=True)
    y = scale(x, k=3.0)
    z = scale(y)
    assert len(made) == 2 and made[0] is not made[1]
    assert made[0].calls == ["forward", "grads"] and made[0].k == 3.0 and made[1].k == 2.0
    assert y.allow_grad and y.op_node is not None and y.op_node.op_name == "Scale"
    z.backward()
    np.testing.assert_array_equal(x.grad.as_numpy(), np.full((2, 3), 6.0, np.float32))
    with md.no_grad():
        w = scale(x)
    assert w.op_node is None and not w.allow_grad and made[-1].calls == ["forward"]
    with pytest.raises(ValueError):
        scale(2.0)
    assert scale.__name__ == "Scale"


def test_linear_relu_second_order(md):
    """allow_higher_order backward through the fused ops: HVP equals the composition's HVP."""
    dims, B = (32, 64, 64, 16), 128
    X_np, Y_np = orc.mlp_data(B, dims[0], dims[-1])
    ps_np = orc.mlp_params(dims)
    vs_np = [np.random.default_rng(100 + i).standard_normal(p.shape).astype(np.float32) for i, p in enumerate(ps_np)]
    want = orc.config5_hvp(X_np, Y_np, ps_np, vs_np)
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    out = fused_mlp(md, X, ps)
    L = ((out - Y) ** 2) / float(out.size)
    L.backward(allow_higher_order=True)
    for i in range(6):
        close(ps[i].grad, want["grads"][i], atol=1e-6)
    s = None
    for p, v in zip(ps, vs_np):
        t = md.sum(p.grad * md.Tensor(v))
        s = t if s is None else s + t
    s.backward()
    for i in range(6):
        close(ps[i].grad, want["hv"][i], rtol=1e-3, atol=1e-6)


def test_linear_relu_shared_input_and_fanout(md):
    """X feeding two fused layers and a fused output consumed twice: gradients accumulate like the
    composition (the epilogue-mask shortcut must only be taken for single-consumer outputs)."""
    relu, linear_relu, linear = fused_ops(md)
    rng = np.random.default_rng(11)
    B, D, H = 512, 256, 384
    X_np = rng.standard_normal((B, D)).astype(np.float32)
    W1_np, W2_np = (rng.standard_normal(s).astype(np.float32) / 16 for s in ((D, H), (H, H)))
    b1_np, b2_np = (rng.standard_normal(H).astype(np.float32) for _ in range(2))

    def build(fused):
        X = md.Tensor(X_np, allow_grad=True)
        W1, W2 = md.Tensor(W1_np, allow_grad=True), md.Tensor(W2_np, allow_grad=True)
        b1, b2 = md.Tensor(b1_np, allow_grad=True), md.Tensor(b2_np, allow_grad=True)
        if fused:
            h = linear_relu(X, W1, b1)
            o = linear_relu(h, W2, b2) + linear_relu(h, W2, b1) + h          # h consumed three times
        else:
            r = lambda t: md.where(t > 0, t, 0)  # noqa: E731
            h = r(X @ W1 + b1)
            o = r(h @ W2 + b2) + r(h @ W2 + b1) + h
        md.sum(o * o).backward()
        return [t.grad.as_numpy() for t in (X, W1, W2, b1, b2)]

    for got, want in zip(build(True), build(False)):
        _close_rms(got, want, "fan-out")
