"""GPU parity of the engine (Tensor / ops / OpNode on device storage) against (a) golden vectors
produced by the REAL reference (tests/golden) and (b) the pinned NumPy oracle on seeded inputs at
sizes the fixtures do not cover.  Tolerances per north_star."""
import os

import numpy as np
import pytest

import np_minidiff as orc
from conftest import GOLDEN, ulp_diff

pytestmark = pytest.mark.gpu


REF_INSTALL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _reference_engine():
    """The UNMODIFIED reference package (offline `pip install --target baseline/_ref /root/reference`,
    git-ignored, shipped to the GPU box) with minidiff_b200.plugin selected through its own
    `--backend` loader: its Tensor / ops / OpNode code then runs on this repo's device backend."""
    import sys

    if not os.path.isdir(os.path.join(REF_INSTALL, "minidiff")):
        pytest.skip("baseline/_ref (offline install of the reference) not present")
    root = os.path.dirname(REF_INSTALL.rstrip(os.sep))
    for p in (os.path.join(os.path.dirname(root), "oracle", "_stubs"), REF_INSTALL):
        if p not in sys.path:
            sys.path.insert(0, p)
    saved = sys.argv
    sys.argv = [saved[0], "--backend", "minidiff_b200.plugin"]
    try:
        import minidiff as ref
    finally:
        sys.argv = saved
    import minidiff_b200.plugin as plugin

    plugin.assert_live(ref)
    return ref


@pytest.fixture(scope="module", params=["b200_engine", "reference_engine"])
def md(request):
    """Every engine test runs twice: with this repo's engine, and with the reference's own engine on
    top of the same device backend (the drop-in boundary of SURVEY 8b, exercised on hardware)."""
    if request.param == "reference_engine":
        return _reference_engine()
    import minidiff_b200 as md

    md.backend.assert_live()
    return md


def _submodule(md, name):
    import importlib

    return importlib.import_module(md.__name__ + "." + name)


def _seed(md, n):
    """device RNG of this repo's backend (the plugin exposes it to the reference engine too)"""
    import minidiff_b200.backend as B

    B.seed(n)


def ours_only(md):
    if md.__name__ != "minidiff_b200":
        pytest.skip("exercises an extension of this repo's engine (not part of the reference API)")


def close(got, want, rtol=1e-4, atol=1e-5):
    got = got.as_numpy() if hasattr(got, "as_numpy") else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def ulps(got, want, n=2):
    got = got.as_numpy()
    assert got.shape == want.shape and got.dtype == want.dtype
    assert ulp_diff(got, want).max(initial=0) <= n


# ------------------------------------------------------------------ C1: README example
def run_c1(md):
    x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
    y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)
    f = 2 * y * md.sin(x) - x**2
    f.backward(allow_higher_order=True)
    out = {"f": f.as_numpy(), "dx": x.grad.as_numpy(), "dy": y.grad.as_numpy()}
    x.grad.backward()
    out["dxx"], out["dxy"] = x.grad.as_numpy(), y.grad.as_numpy()
    return out


def test_c1_readme_first_and_second_order(md, golden):
    g = golden("c1.npz")
    r = run_c1(md)
    # the expression cancels (6*sin(2) - 4 ~ 1.46): a 1-ulp difference in sin() is amplified in
    # ulps of the small result, so the 2-ulp budget is applied at the magnitude of the largest
    # intermediate term (|x**2|, |2*y*sin x| <= 16): atol = 2 * 2^-23 * 16
    for k in ("f", "dx", "dy", "dxx", "dxy"):
        assert r[k].dtype == np.float32 and r[k].shape == (2, 4)
        np.testing.assert_allclose(r[k], g[k], rtol=0, atol=2 * 2.0**-23 * 16, err_msg=k)


def test_c1_default_dtype_follows_numpy(md):
    x = md.Tensor([[0, 2, -2, 1]], allow_grad=True)          # README literal ints -> int64
    assert x.dtype == np.int64
    f = 2 * x * md.sin(x)
    assert f.dtype == np.float64


# ------------------------------------------------------------------ C2: broadcast chain
def run_c2(md, a_np, c_np, higher=False):
    a = md.Tensor(a_np.copy(), allow_grad=True)
    c = md.Tensor(c_np.copy(), allow_grad=True)
    z = md.sin(a * c + a) ** 2
    loss = md.sum(z)
    loss.backward(allow_higher_order=higher) if higher else loss.backward()
    return loss, a, c


@pytest.mark.parametrize("n,m", [(64, 48), (257, 131), (1024, 1024)])
def test_c2_vs_golden(md, golden, n, m):
    g = golden(f"c2_{n}x{m}.npz")
    loss, a, c = run_c2(md, g["a"], g["c"])
    assert loss.shape == () and a.grad.shape == (n, 1) and c.grad.shape == (1, m)
    close(loss, g["loss"], rtol=1e-4)
    close(a.grad, g["da"], rtol=1e-4, atol=1e-5 * np.sqrt(m))
    close(c.grad, g["dc"], rtol=1e-4, atol=1e-5 * np.sqrt(n))


def test_c2_unfused_reference_chain_matches_fused(md, golden):
    """allow_higher_order=True disables the fused handlers (every grad op is recorded, like the
    reference): both routes must agree."""
    g = golden("c2_257x131.npz")
    _, a1, c1 = run_c2(md, g["a"], g["c"], higher=False)
    _, a2, c2 = run_c2(md, g["a"], g["c"], higher=True)
    close(a1.grad, a2.grad.as_numpy(), rtol=1e-5, atol=1e-4)
    close(c1.grad, c2.grad.as_numpy(), rtol=1e-5, atol=1e-4)


def test_c2_vs_oracle_medium(md):
    a_np, c_np = orc.config2_inputs(2048, 3072)
    want = orc.config2(a_np, c_np)
    loss, a, c = run_c2(md, a_np, c_np)
    truth_scale = np.sqrt(3072)
    close(loss, want["loss"], rtol=1e-4)
    close(a.grad, want["da"], rtol=1e-4, atol=2e-5 * truth_scale)
    close(c.grad, want["dc"], rtol=1e-4, atol=2e-5 * truth_scale)


def test_c2_full_size_properties(md):
    """BASELINE size (2^26 elements): size-independent properties instead of a CPU recomputation:
    linearity of the gradient in the upstream seed and agreement of loss with a two-block split."""
    n = m = 8192
    a_np, c_np = orc.config2_inputs(n, m)
    loss, a, c = run_c2(md, a_np, c_np)
    l_full = float(loss.item())
    parts = 0.0
    for lo, hi in ((0, n // 2), (n // 2, n)):
        l, _, _ = run_c2(md, a_np[lo:hi], c_np)
        parts += float(l.item())
    assert abs(l_full - parts) <= 1e-5 * abs(l_full)
    # row block gradients of `a` must equal the corresponding rows of the full gradient
    _, a_half, _ = run_c2(md, a_np[: n // 2], c_np)
    np.testing.assert_allclose(a.grad.as_numpy()[: n // 2], a_half.grad.as_numpy(), rtol=1e-5, atol=1e-3)
    # oracle on a row sample (finishes in well under a second)
    want = orc.config2(a_np[:64], c_np)
    np.testing.assert_allclose(a.grad.as_numpy()[:64], want["da"], rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ C3: matmul fwd + both grads
@pytest.mark.parametrize("shape", ["48x40x56", "256x384x128"])
def test_c3_vs_golden(md, golden, shape):
    g = golden(f"c3_{shape}.npz")
    A, B = md.Tensor(g["A"], allow_grad=True), md.Tensor(g["B"], allow_grad=True)
    C = A @ B
    C.backward()
    K = g["A"].shape[1]
    close(C, g["C"], atol=1e-5 * np.sqrt(K))
    close(A.grad, g["dA"], atol=1e-5 * np.sqrt(g["B"].shape[1]))
    close(B.grad, g["dB"], atol=1e-5 * np.sqrt(g["A"].shape[0]))


@pytest.mark.parametrize("n", [1024, 2048])
def test_c3_vs_float64(md, n):
    A_np = np.random.default_rng(1234).standard_normal((n, n)).astype(np.float32)
    B_np = np.random.default_rng(1235).standard_normal((n, n)).astype(np.float32)
    A, B = md.Tensor(A_np, allow_grad=True), md.Tensor(B_np, allow_grad=True)
    C = A @ B
    C.backward()
    A64, B64 = A_np.astype(np.float64), B_np.astype(np.float64)
    ones = np.ones((n, n))
    tol = dict(rtol=1e-4, atol=1e-5 * np.sqrt(n))
    np.testing.assert_allclose(C.as_numpy(), A64 @ B64, **tol)
    np.testing.assert_allclose(A.grad.as_numpy(), ones @ B64.T, **tol)
    np.testing.assert_allclose(B.grad.as_numpy(), A64.T @ ones, **tol)


def test_c3_full_size_properties(md):
    """BASELINE size (8192^3, all three GEMMs on the CTA-pair tcgen05 kernel): no CPU recomputation
    of 1.1 TFLOP; instead (i) a random sample of C entries against float64 dot products,
    (ii) dA = 1 @ B^T means every row of dA equals the row sums of B, and dB = A^T @ 1 means every
    column of dB equals the column sums of A -- checked against float64 sums."""
    n = 8192
    A_np = np.random.default_rng(1234).standard_normal((n, n), dtype=np.float32)
    B_np = np.random.default_rng(1235).standard_normal((n, n), dtype=np.float32)
    A, B = md.Tensor(A_np, allow_grad=True), md.Tensor(B_np, allow_grad=True)
    C = A @ B
    C.backward()
    Cn = C.as_numpy()
    rng = np.random.default_rng(0)
    ii, jj = rng.integers(0, n, 256), rng.integers(0, n, 256)
    truth = np.einsum("sk,ks->s", A_np[ii].astype(np.float64), B_np[:, jj].astype(np.float64))
    np.testing.assert_allclose(Cn[ii, jj], truth, rtol=1e-4, atol=1e-5 * np.sqrt(n))
    # a full tile row / column too (catches a wrong tile mapping that a sparse sample could miss)
    np.testing.assert_allclose(Cn[4097], A_np[4097].astype(np.float64) @ B_np.astype(np.float64),
                               rtol=1e-4, atol=1e-5 * np.sqrt(n))
    row_sums_B = B_np.astype(np.float64).sum(axis=1)          # dA[i, k] = sum_j B[k, j]
    col_sums_A = A_np.astype(np.float64).sum(axis=0)          # dB[k, j] = sum_i A[i, k]
    dA, dB = A.grad.as_numpy(), B.grad.as_numpy()
    tol = dict(rtol=1e-4, atol=1e-5 * np.sqrt(n))
    for i in (0, 127, 128, 4095, 8191):
        np.testing.assert_allclose(dA[i], row_sums_B, **tol)
    for j in (0, 255, 256, 5000, 8191):
        np.testing.assert_allclose(dB[:, j], col_sums_A, **tol)
    assert np.ptp(dA, axis=0).max() <= 2e-5 * np.sqrt(n) and np.ptp(dB, axis=1).max() <= 2e-5 * np.sqrt(n)


def test_c4_full_size_properties(md):
    """BASELINE size (batch 65536, 1024-4096-4096-1024): the mean-MSE gradient of the full batch is
    the average of the gradients of its two halves (the identity data parallelism relies on), and
    the last bias gradient equals 2/(B*1024) * column sums of (out - Y) recomputed with device ops."""
    from minidiff_b200 import workloads as W

    B = 65536
    X_np, Y_np = W.mlp_data(B, 1024, 1024, seed=3)
    init = W.mlp_params()

    def grads_of(Xs, Ys):
        ps = [md.Tensor(p.copy(), allow_grad=True) for p in init]
        X, Y = md.Tensor(Xs), md.Tensor(Ys)
        out = mlp(md, X, ps)
        loss = md.mean((out - Y) ** 2)
        loss.backward()
        resid = md.sum(out - Y, axis=0).as_numpy().astype(np.float64)
        return float(loss.item()), [p.grad.as_numpy() for p in ps], resid

    l_full, g_full, resid = grads_of(X_np, Y_np)
    np.testing.assert_allclose(g_full[5], 2.0 * resid / (B * 1024.0), rtol=1e-4, atol=1e-9)
    l_a, g_a, _ = grads_of(X_np[: B // 2], Y_np[: B // 2])
    l_b, g_b, _ = grads_of(X_np[B // 2:], Y_np[B // 2:])
    assert abs(l_full - 0.5 * (l_a + l_b)) <= 1e-5 * abs(l_full)
    for gf, ga, gb in zip(g_full, g_a, g_b):
        scale = np.abs(gf).max()
        np.testing.assert_allclose(gf, 0.5 * (ga + gb), rtol=1e-4, atol=1e-4 * scale)


# ------------------------------------------------------------------ C4: MLP training step
DIMS, BATCH = (16, 32, 32, 8), 64


def relu(md, h):
    return md.where(h > 0, h, 0)


def mlp(md, X, ps):
    h = X
    n = len(ps) // 2
    for l in range(n):
        h = h @ ps[2 * l] + ps[2 * l + 1]
        if l < n - 1:
            h = relu(md, h)
    return h


def run_c4(md, X_np, Y_np, ps_np, lr=0.01):
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    loss = md.mean((mlp(md, X, ps) - Y) ** 2)
    loss.backward()
    grads = [p.grad.as_numpy() for p in ps]
    with md.no_grad():
        for p in ps:
            p -= lr * p.grad
    return loss, grads, ps


def test_c4_vs_golden(md, golden):
    g = golden("c4.npz")
    X, Y = orc.mlp_data(BATCH, DIMS[0], DIMS[-1])
    loss, grads, ps = run_c4(md, X, Y, orc.mlp_params(DIMS))
    close(loss, g["loss"])
    for i in range(6):
        close(grads[i], g[f"g{i}"])
        close(ps[i], g[f"p{i}"])


def test_c4_vs_oracle_wide(md):
    dims, B = (256, 512, 512, 128), 1024
    X, Y = orc.mlp_data(B, dims[0], dims[-1])
    ps_np = orc.mlp_params(dims)
    want = orc.config4_step(X, Y, ps_np)
    loss, grads, ps = run_c4(md, X, Y, ps_np)
    close(loss, want["loss"])
    for i in range(6):
        close(grads[i], want["grads"][i], atol=1e-6)
        close(ps[i], want["params"][i], atol=1e-6)
    # in-place update did not re-allocate parameter storage, graph refs behave like the reference
    assert all(p.grad is not None for p in ps)


BASELINE_DIMS = (1024, 4096, 4096, 1024)


def _gemm_paths(reset=False):
    import ctypes as C

    from minidiff_b200.backend._lib import check, lib

    counts = (C.c_uint64 * 8)()
    check(lib.mdb_gemm_stats(counts, 1 if reset else 0))
    return dict(simt=int(counts[0]), single=int(counts[1]), presplit=int(counts[2]), pair=int(counts[3]),
                pair_streamk=int(counts[4]))


def _close_rms(got, want, what, atol=1e-5):
    """north_star tolerance for reductions / GEMMs, rtol 1e-4 / atol 1e-5, with the atol scaled by
    the rms of the reference values (the gradients here are ~1e-5..1e-2, not ~1)."""
    want = np.asarray(want)
    rms = float(np.sqrt(np.mean(np.square(want.astype(np.float64)))))
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=atol * rms, err_msg=what)


_CASES = {}


def _c4_baseline_case():
    """oracle + float64 results of the BASELINE-dims C4 step, computed once for both engines"""
    if "c4" not in _CASES:
        X, Y, ps_np = orc.kink_safe_mlp_batch(2048, BASELINE_DIMS)
        want = orc.config4_step(X, Y, ps_np)
        truth = orc.config4_step(X.astype(np.float64), Y.astype(np.float64), [p.astype(np.float64) for p in ps_np])
        _CASES["c4"] = (X, Y, ps_np, want, truth)
    return _CASES["c4"]


def _c5_baseline_case():
    if "c5" not in _CASES:
        X, Y, ps_np = orc.kink_safe_mlp_batch(8192, BASELINE_DIMS)
        vs_np = [np.random.default_rng(100 + i).standard_normal(p.shape).astype(np.float32)
                 for i, p in enumerate(ps_np)]
        want = orc.config5_hvp(X, Y, ps_np, vs_np)
        truth = orc.config5_hvp(X.astype(np.float64), Y.astype(np.float64), [p.astype(np.float64) for p in ps_np],
                                [v.astype(np.float64) for v in vs_np])
        _CASES["c5"] = (X, Y, ps_np, vs_np, want, truth)
    return _CASES["c5"]


def test_c4_baseline_dims_vs_oracle_on_the_pair_kernel(md):
    """One full C4 training step at the BASELINE layer dims (1024-4096-4096-1024) on a 2048-row batch
    against the oracle AND float64, rtol 1e-4 / atol 1e-5*rms.  Every GEMM of the step (3 forward,
    3 dW, 2 dh) must have run on the CTA-pair tcgen05 kernel the benchmark measures."""
    from minidiff_b200.backend._lib import check, lib

    X, Y, ps_np, want, truth = _c4_baseline_case()
    check(lib.mdb_gemm_config(2))
    check(lib.mdb_gemm_tune(4 | 32))       # the pair kernel wherever it is legal (it is for all 8 GEMMs here)
    _gemm_paths(reset=True)
    try:
        loss, grads, ps = run_c4(md, X, Y, ps_np)
    finally:
        check(lib.mdb_gemm_tune(4))
        check(lib.mdb_gemm_config(0))
    paths = _gemm_paths()
    assert paths["pair"] + paths["pair_streamk"] == 8 and paths["simt"] == paths["single"] == paths["presplit"] == 0, paths
    close(loss, want["loss"], rtol=1e-5)
    for i in range(6):
        _close_rms(grads[i], want["grads"][i], f"grad {i} vs oracle")
        _close_rms(grads[i], truth["grads"][i], f"grad {i} vs float64")
        _close_rms(ps[i].as_numpy(), want["params"][i], f"param {i} vs oracle")


def test_c4_baseline_dims_with_the_fast_gemm_split(md):
    """Same step with the opt-in "fast" operand split (one TF32 MMA + two BF16 cross-term MMAs per product,
    backend.set_matmul_split): against the oracle and float64 at rtol 1e-4 / atol 3e-5*rms -- three times the
    default's atol, which is why the split is opt-in -- and the gradients' rms error stays below 4e-6 of their rms."""
    from minidiff_b200.backend._lib import check, lib

    X, Y, ps_np, want, truth = _c4_baseline_case()
    check(lib.mdb_gemm_config(2))
    check(lib.mdb_gemm_tune(4 | 32))
    import minidiff_b200.backend as device_backend      # process-wide switch of the C-ABI library, either engine

    device_backend.set_matmul_split("fast")
    _gemm_paths(reset=True)
    try:
        loss, grads, ps = run_c4(md, X, Y, ps_np)
    finally:
        device_backend.set_matmul_split("3xtf32")
        check(lib.mdb_gemm_tune(4))
        check(lib.mdb_gemm_config(0))
    paths = _gemm_paths()
    assert paths["pair"] + paths["pair_streamk"] == 8, paths
    close(loss, want["loss"], rtol=1e-5)
    for i in range(6):
        _close_rms(grads[i], want["grads"][i], f"grad {i} vs oracle", atol=3e-5)
        _close_rms(grads[i], truth["grads"][i], f"grad {i} vs float64", atol=3e-5)
        t = np.asarray(truth["grads"][i], dtype=np.float64)
        assert np.sqrt(np.mean((grads[i] - t) ** 2)) < 4e-6 * np.sqrt(np.mean(t ** 2)), i


def test_c5_full_baseline_config_vs_oracle_and_float64(md):
    """The FULL BASELINE config 5 (batch 8192, 1024-4096-4096-1024): first-order gradients of the
    unreduced loss and the Hessian-vector product against the oracle and against float64, rtol 1e-4 /
    atol 1e-5*rms; all 22 GEMMs (F-order / C-of-F-order / stride-0 operand mix of the second-order
    graph) on the tensor-core path, none on the CUDA-core kernel."""
    from minidiff_b200.backend._lib import check, lib

    X, Y, ps_np, vs_np, want, truth = _c5_baseline_case()
    check(lib.mdb_gemm_config(2))
    _gemm_paths(reset=True)
    try:
        g1, hv = run_c5(md, X, Y, ps_np, vs_np)
    finally:
        check(lib.mdb_gemm_config(0))
    paths = _gemm_paths()
    assert paths["simt"] == 0 and sum(paths.values()) == 22, paths
    for i in range(6):
        _close_rms(g1[i], want["grads"][i], f"grad {i} vs oracle")
        _close_rms(hv[i], want["hv"][i], f"hv {i} vs oracle")
        _close_rms(hv[i], truth["hv"][i], f"hv {i} vs float64")


def test_c4_param_update_requires_no_grad(md):
    X, Y = orc.mlp_data(8, DIMS[0], DIMS[-1])
    ps = [md.Tensor(p, allow_grad=True) for p in orc.mlp_params(DIMS)]
    loss = md.mean((mlp(md, md.Tensor(X), ps) - md.Tensor(Y)) ** 2)
    loss.backward()
    with pytest.raises(ValueError):
        ps[0] -= 0.01 * ps[0].grad       # reference: graph_refs never drop on leaves


# ------------------------------------------------------------------ C5: Hessian-vector product
def run_c5(md, X_np, Y_np, ps_np, vs_np):
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    vs = [md.Tensor(v) for v in vs_np]
    out = mlp(md, X, ps)
    L = ((out - Y) ** 2) / float(out.size)
    L.backward(allow_higher_order=True)
    g1 = [p.grad.as_numpy() for p in ps]
    s = None
    for p, v in zip(ps, vs):
        t = md.sum(p.grad * v)
        s = t if s is None else s + t
    s.backward()
    return g1, [p.grad.as_numpy() for p in ps]


def test_c5_hvp_vs_golden(md, golden):
    g = golden("c5.npz")
    X, Y = orc.mlp_data(BATCH, DIMS[0], DIMS[-1])
    g1, hv = run_c5(md, X, Y, orc.mlp_params(DIMS), [g[f"v{i}"] for i in range(6)])
    for i in range(6):
        close(g1[i], g[f"g{i}"], atol=1e-6)
        close(hv[i], g[f"hv{i}"], atol=1e-6)


def test_c5_hvp_vs_oracle_and_float64(md):
    dims, B = (64, 128, 128, 32), 256
    X, Y = orc.mlp_data(B, dims[0], dims[-1])
    ps_np = orc.mlp_params(dims)
    vs_np = [np.random.default_rng(100 + i).standard_normal(p.shape).astype(np.float32)
             for i, p in enumerate(ps_np)]
    want = orc.config5_hvp(X, Y, ps_np, vs_np)
    truth = orc.config5_hvp(X.astype(np.float64), Y.astype(np.float64),
                            [p.astype(np.float64) for p in ps_np],
                            [v.astype(np.float64) for v in vs_np])
    g1, hv = run_c5(md, X, Y, ps_np, vs_np)
    for i in range(6):
        close(hv[i], want["hv"][i], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(hv[i], truth["hv"][i], rtol=1e-3, atol=1e-6)


def test_scalar_loss_second_order_raises_like_reference(md):
    """SURVEY finding 1: reducing to a scalar before a higher-order backward raises in the
    reference (unbroadcast's gradient is wrong for up-broadcasts); same behaviour here."""
    x = md.Tensor(np.arange(6, dtype=np.float32).reshape(2, 3), allow_grad=True)
    loss = md.sum(x ** 3)
    loss.backward(allow_higher_order=True)
    with pytest.raises(ValueError):
        md.sum(x.grad * x.grad).backward()


# ------------------------------------------------------------------ per-op golden vectors
_D = np.load(os.path.join(GOLDEN, "ops.npz"))
_NAMES = sorted({k.split("/")[0] for k in _D.files})


def _funcs(md):
    F = {
        "neg": lambda t: -t, "where_relu": lambda t: md.where(t > 0, t, 0),
        "where_tt": lambda t, u: md.where(t > u, t, u), "clip": lambda t: md.clip(t, -0.5, 0.5),
        "clip_lo": lambda t: md.clip(t, 0, None), "mask_mul": lambda t: t * (t > 0),
        "sum_all": md.sum, "mean_all": md.mean, "max_all": md.max,
        "sum_ax0": lambda t: md.sum(t, axis=(0,)),
        "sum_ax1_keep": lambda t: md.sum(t, axis=(1,), keepdims=True),
        "sum_ax02": lambda t: md.sum(t, axis=(0, 2)), "max_ax1": lambda t: md.max(t, axis=1),
        "min_ax1": lambda t: md.min(t, axis=1), "prod_ax0": lambda t: md.prod(t, axis=0),
        "T_matmul": lambda t, u: t.T @ u, "reshape": lambda t: md.reshape(t, (7, 5)),
        "broadcast_to": lambda t: md.broadcast_to(t, (4, 5, 7)),
        "expand_dims": lambda t: md.expand_dims(t, 1), "swapaxes": lambda t: md.swapaxes(t, 0, 2),
        "flip": lambda t: md.flip(t, axis=1), "getitem_slice": lambda t: t[1:4, ::2],
        "getitem_int": lambda t: t[2], "tensordot": lambda t, u: md.tensordot(t, u, axes=1),
        "mod": lambda t: md.mod(t, 0.75), "std_ax1": lambda t: md.std(t, axis=(1,)),
        "chain_bcast": lambda a, c: md.sin(a * c + a) ** 2, "power_tt": md.power,
    }
    for n in ("add", "subtract", "multiply", "true_divide"):
        f = getattr(md, n)
        F[n + "_bcast_col_row"] = f
        F[n + "_bcast_vec"] = f
        F[n + "_scalar_l"] = (lambda f: lambda t: f(2.5, t))(f)
        F[n + "_scalar_r"] = (lambda f: lambda t: f(t, 2.5))(f)
    for e in (2, 1, 0, 0.5, -1, 3, 2.5):
        F[f"power_{e}"] = (lambda e: lambda t: t ** e)(e)
    return F


@pytest.mark.parametrize("name", _NAMES)
def test_op_vectors_forward_and_backward(md, name):
    fn = _funcs(md).get(name) or getattr(md, name)
    ins, i = [], 0
    while f"{name}/in{i}" in _D.files:
        ins.append(_D[f"{name}/in{i}"]); i += 1
    ts = [md.Tensor(v.copy(), allow_grad=True) for v in ins]
    out = fn(*ts)
    want = _D[f"{name}/out"]
    got = out.as_numpy()
    assert got.shape == want.shape and got.dtype == want.dtype, (got.shape, want.shape, got.dtype, want.dtype)
    if want.dtype == np.float32:
        loose = name in ("matmul", "T_matmul", "tensordot", "dot", "std_ax1", "prod_ax0") or name.startswith(("sum", "mean"))
        if loose:
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
        else:
            assert ulp_diff(got, want).max(initial=0) <= 2, ulp_diff(got, want).max()
    else:
        np.testing.assert_array_equal(got, want)
    if out.allow_grad and not out.is_leaf:
        w = md.Tensor(np.random.default_rng(11).standard_normal(out.shape).astype(out.dtype))
        if f"{name}/raises" in _D.files:
            with pytest.raises(Exception) as ei:
                md.sum(out * w).backward()
            assert type(ei.value).__name__ == str(_D[f"{name}/raises"]), ei.value
            return
        md.sum(out * w).backward()
        for j, t in enumerate(ts):
            if f"{name}/g{j}" in _D.files:
                gw = _D[f"{name}/g{j}"]
                gg = t.grad.as_numpy()
                assert gg.shape == gw.shape and gg.dtype == gw.dtype, (name, j, gg.dtype, gw.dtype)
                np.testing.assert_allclose(gg, gw, rtol=1e-4, atol=1e-5)
            else:
                assert t.grad is None


# ------------------------------------------------------------------ finite-difference checks
FD_OPS = ["sin", "cos", "exp", "tanh", "sinh", "cosh", "absolute", "copy"]


@pytest.mark.parametrize("name", FD_OPS)
def test_first_order_fd_through_utils(md, name):
    """reference tests/test_ops.py:25-62 harness: loss = sum((0 - f(x))**2)/2, h = 1e-2,
    rtol 1e-3 / atol 1e-4, float64 inputs (NumPy-default dtype of md.randn)."""
    compute_grads = _submodule(md, "utils").compute_grads
    _seed(md, 3)
    x = md.randn(2, 2, 2, 2, allow_grad=True)
    op = getattr(md, name)

    def loss(t):
        actual = op(t)
        return md.sum((md.zeros_like(actual) - actual) ** 2) / 2

    manual, auto = compute_grads(x, func=loss, h=1e-2)
    np.testing.assert_allclose(manual[0].as_numpy(), auto[0].as_numpy(), rtol=1e-3, atol=1e-4)


def test_fd_binary_broadcast_and_matmul(md):
    compute_grads = _submodule(md, "utils").compute_grads
    _seed(md, 5)
    a, b = md.randn(3, 1, allow_grad=True), md.randn(1, 4, allow_grad=True)
    manual, auto = compute_grads(a, b, func=lambda p, q: md.sum(md.sin(p * q + p) ** 2), h=1e-3)
    for m_, a_ in zip(manual, auto):
        np.testing.assert_allclose(m_.as_numpy(), a_.as_numpy(), rtol=1e-3, atol=1e-4)
    x, y = md.randn(5, 6, allow_grad=True), md.randn(6, 4, allow_grad=True)
    manual, auto = compute_grads(x, y, func=lambda p, q: md.sum((p @ q) ** 2) / 2, h=1e-3)
    for m_, a_ in zip(manual, auto):
        np.testing.assert_allclose(m_.as_numpy(), a_.as_numpy(), rtol=1e-3, atol=1e-4)


def test_second_order_fd(md):
    """d/dx of the autodiff gradient, checked by central differences of the first-order gradient."""
    x0 = np.random.default_rng(2).standard_normal((3, 4))

    def grad_at(v):
        x = md.Tensor(v, allow_grad=True)
        f = md.sin(x) * x ** 3
        f.backward()
        return x.grad.as_numpy()

    x = md.Tensor(x0, allow_grad=True)
    f = md.sin(x) * x ** 3
    f.backward(allow_higher_order=True)
    x.grad.backward()
    h = 1e-5
    fd = (grad_at(x0 + h) - grad_at(x0 - h)) / (2 * h)       # f is elementwise: Hessian is diagonal
    np.testing.assert_allclose(x.grad.as_numpy(), fd, rtol=1e-6, atol=1e-6)


def test_grad_aliasing_and_accumulation_semantics(md):
    """SURVEY finding 3: identity gradients alias; fused in-place accumulation must not corrupt
    aliased buffers nor a user's saved reference."""
    a = md.Tensor(np.ones((4, 4), np.float32), allow_grad=True)
    b = md.Tensor(np.full((4, 4), 2, np.float32), allow_grad=True)
    f = a + b
    f.backward(retain_grads=True)
    assert a.grad is b.grad is f.grad
    g = (a * b + a * a + a)
    g.backward()
    np.testing.assert_array_equal(a.grad.as_numpy(), np.full((4, 4), 2 + 2 + 1, np.float32))
    np.testing.assert_array_equal(b.grad.as_numpy(), np.ones((4, 4), np.float32))
    keep = a.grad
    snapshot = keep.as_numpy().copy()
    (a * 3).backward(reset_grads=False)                     # accumulates on top
    np.testing.assert_array_equal(a.grad.as_numpy(), snapshot + 3)
    np.testing.assert_array_equal(keep.as_numpy(), snapshot)  # user's reference untouched
    s = md.sum(a)
    s.backward()
    assert a.grad._data.strides == (0, 0) and not a.grad._data.writeable


def test_reuse_graph_cache_gives_same_grads(md):
    x_np = np.random.default_rng(0).standard_normal((6, 5)).astype(np.float32)

    def step(x):
        y = md.sum(md.sin(x) * x + x)
        y.backward()
        return x.grad.as_numpy()

    plain = step(md.Tensor(x_np, allow_grad=True))
    with _submodule(md, "caching").reuse_graph():
        c1 = step(md.Tensor(x_np, allow_grad=True))
        c2 = step(md.Tensor(x_np, allow_grad=True))
    np.testing.assert_array_equal(plain, c1)
    np.testing.assert_array_equal(plain, c2)


def test_custom_op_via_create_op_func(md):
    """README 'Custom Functions': as_minidiff + create_op_func on a backend function."""
    B = md.backend
    wr = _submodule(md, "ops.wrapping")       # the reference keeps the factories in ops/wrapping.py
    softsign = wr.create_unary_op_func(
        forward_func=wr.as_minidiff(lambda a: B.true_divide(a, B.add(B.absolute(a), 1))),
        grad=lambda x, grad: grad / (md.absolute(x) + 1) ** 2, op_name="softsign")
    x_np = np.linspace(-2, 2, 9, dtype=np.float32)
    x = md.Tensor(x_np, allow_grad=True)
    y = softsign(x)
    y.backward()
    np.testing.assert_allclose(y.as_numpy(), x_np / (np.abs(x_np) + 1), rtol=1e-6)
    np.testing.assert_allclose(x.grad.as_numpy(), 1 / (np.abs(x_np) + 1) ** 2, rtol=1e-6)


def test_host_batch_feeder_pipeline_matches_resident_inputs(md):
    ours_only(md)
    """e2e input pipeline (pinned host -> copy stream -> double-buffered device inputs) must feed
    exactly the bytes given, step after step, while uploads overlap compute."""
    from minidiff_b200 import workloads as W

    dims, B = (64, 128, 128, 32), 512
    X, Y = orc.mlp_data(B, dims[0], dims[-1])
    ref = [md.Tensor(p.copy(), allow_grad=True) for p in orc.mlp_params(dims)]
    got = [md.Tensor(p.copy(), allow_grad=True) for p in orc.mlp_params(dims)]
    feeder = W.HostBatchFeeder(X, Y)
    Xr, Yr = md.Tensor(X), md.Tensor(Y)
    for _ in range(4):
        l_ref = W.mlp_train_step(Xr, Yr, ref, 0.05)
        Xd, Yd = feeder.next()
        l_got = W.mlp_train_step(Xd, Yd, got, 0.05)
        assert float(l_ref.item()) == float(l_got.item())
    for a, b in zip(ref, got):
        np.testing.assert_array_equal(a.as_numpy(), b.as_numpy())


# ---------------------------------------------------------------- CUDA-graph capture
def test_captured_graph_replays_c1_with_refreshed_inputs(md):
    ours_only(md)
    """Capture forward + first-order backward of the README expression once, then replay it on new
    input values written IN PLACE: results must equal an eager evaluation bit for bit."""
    rng = np.random.default_rng(3)
    x = md.Tensor(rng.standard_normal((2, 4)).astype(np.float32), allow_grad=True)
    y = md.Tensor(rng.standard_normal((2, 4)).astype(np.float32), allow_grad=True)

    def step():
        f = 2 * y * md.sin(x) - x ** 2
        f.backward()
        return f

    g = md.capture_graph(step)
    assert g.kernel_launches >= 5
    for seed in (10, 11):
        xn = np.random.default_rng(seed).standard_normal((2, 4)).astype(np.float32)
        yn = np.random.default_rng(seed + 100).standard_normal((2, 4)).astype(np.float32)
        with md.no_grad():               # leaves that already sit in a graph refuse in-place writes
            x[...] = xn                  # (reference tensor.py:257-264) unless grad tracking is off
            y[...] = yn
        f = g.replay()
        got = (f.as_numpy().copy(), x.grad.as_numpy().copy(), y.grad.as_numpy().copy())
        xe, ye = md.Tensor(xn, allow_grad=True), md.Tensor(yn, allow_grad=True)
        fe = 2 * ye * md.sin(xe) - xe ** 2
        fe.backward()
        np.testing.assert_array_equal(got[0], fe.as_numpy())
        np.testing.assert_array_equal(got[1], xe.grad.as_numpy())
        np.testing.assert_array_equal(got[2], ye.grad.as_numpy())
    g.close()


def test_captured_training_step_matches_eager_and_pins_its_memory(md):
    ours_only(md)
    """K replays of a captured MLP training step == K eager steps; memory allocated by OTHER work
    between replays never aliases the graph's buffers (private pool)."""
    from minidiff_b200 import workloads as W

    dims, B, K = (64, 128, 128, 32), 256, 4
    X_np, Y_np = W.mlp_data(B, dims[0], dims[-1], seed=5)
    init = W.mlp_params(dims)
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    eager = [md.Tensor(p.copy(), allow_grad=True) for p in init]
    # warmup=0 so that the captured parameters see exactly K updates
    graph_params = [md.Tensor(p.copy(), allow_grad=True) for p in init]
    W.mlp_train_step(X, Y, [md.Tensor(p.copy(), allow_grad=True) for p in init])   # first-launch setup
    g = md.capture_graph(lambda: W.mlp_train_step(X, Y, graph_params), warmup=0)
    losses = []
    for _ in range(K):
        loss = g.replay()
        junk = [md.backend.ones((257, 129)) * 3.0 for _ in range(8)]   # churn the general allocator
        del junk
        losses.append(float(loss.item()))
    want = [float(W.mlp_train_step(X, Y, eager).item()) for _ in range(K)]
    np.testing.assert_array_equal(np.float32(losses), np.float32(want))
    for a, b in zip(graph_params, eager):
        np.testing.assert_array_equal(a.as_numpy(), b.as_numpy())
    assert g.pinned_bytes > 0
    g.close()


def test_capture_refuses_readbacks(md):
    ours_only(md)
    x = md.Tensor(np.ones((4, 4), np.float32), allow_grad=True)

    def bad():
        return float(md.sum(x * x).item())        # a read-back inside the capture

    with pytest.raises(RuntimeError):
        md.capture_graph(bad, warmup=1)
    # the library is usable again afterwards
    np.testing.assert_allclose(md.sum(x * x).item(), 16.0)


def test_training_steps_are_bitwise_reproducible(md):
    """Three C4 training steps at the BASELINE layer dims from the same initial state, run twice: every
    parameter bit-identical (the GEMM's stream-K fix-up adds partial tiles in cluster order, split
    reductions fold their partials in a fixed order, nothing uses floating-point atomics)."""
    X_np, Y_np, ps_np, _, _ = _c4_baseline_case()

    def run():
        X, Y = md.Tensor(X_np), md.Tensor(Y_np)
        ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
        for _ in range(3):
            loss = md.mean((mlp(md, X, ps) - Y) ** 2)
            loss.backward()
            with md.no_grad():
                for p in ps:
                    p -= 0.01 * p.grad
        return [p.as_numpy() for p in ps], loss.as_numpy()

    (a, la), (b, lb) = run(), run()
    assert la.tobytes() == lb.tobytes()
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
