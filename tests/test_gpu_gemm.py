"""tcgen05 3xTF32 GEMM parity (forced tensor-core path) for every operand-major combination the
engine produces: NN (forward), NT (dA = dC @ B^T), TN (dB = A^T @ dC), TT; ragged extents (TMA
zero-fill + predicated epilogue), accumulate-into-C, and the fp32-level accuracy 3xTF32 must keep
(north_star: rtol 1e-4 / atol 1e-5 against NumPy fp32; both are also compared with float64)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import minidiff_b200.backend as backend

    backend.assert_live()
    return backend


@pytest.fixture(params=["auto", "pair", "single"])
def force_tc(B, request):
    """tensor-core kernels only (no silent CUDA-core fallback).  RAW mode is taken when TMA can
    address the operands in place; ragged pitches / doubly strided views take PRE-SPLIT mode.
    Every case runs three times: dispatcher's choice, the CTA-pair kernel (cta_group::2, 256x256
    tiles) wherever it is legal, and the single-CTA kernel (128x128 tiles) only."""
    from minidiff_b200.backend._lib import check, lib

    check(lib.mdb_gemm_config(2))
    check(lib.mdb_gemm_tune({"auto": 4, "pair": 4 | 32, "single": 4 | 16}[request.param]))
    yield
    check(lib.mdb_gemm_tune(4))
    check(lib.mdb_gemm_config(0))


def operands(B, M, K, N, layout, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    da = B.asarray(a) if layout[0] == "N" else B.asarray(np.ascontiguousarray(a.T)).T
    db = B.asarray(b) if layout[1] == "N" else B.asarray(np.ascontiguousarray(b.T)).T
    return a, b, da, db


SHAPES = [(128, 128, 128), (128, 160, 128), (256, 512, 384), (129, 200, 130), (1000, 777, 555),
          (384, 4096, 64), (64, 8192, 64), (2048, 1024, 1536), (300, 260, 272), (512, 96, 768),
          (4352, 512, 2304)]


@pytest.mark.parametrize("layout", ["NN", "NT", "TN", "TT"])
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_forced_tensor_core_path(B, force_tc, M, K, N, layout):
    a, b, da, db = operands(B, M, K, N, layout, seed=M + K + N)
    got = B.matmul(da, db).numpy()
    truth = a.astype(np.float64) @ b.astype(np.float64)
    ref32 = a @ b
    scale = np.sqrt(K)
    err = np.abs(got - truth)
    err32 = np.abs(ref32 - truth)
    # 3xTF32 keeps fp32-class accuracy: within the budget against float64 truth, and no worse than
    # a small multiple of NumPy/OpenBLAS's own fp32 error
    assert err.max() <= 1e-5 * scale + 1e-4 * np.abs(truth).max() * 1e-2, (err.max(), err32.max())
    assert err.max() <= 8 * err32.max() + 1e-6 * scale, (err.max(), err32.max())
    np.testing.assert_allclose(got, ref32, rtol=1e-4, atol=1e-5 * scale)


def test_plain_tf32_would_fail_the_budget(B, force_tc):
    """Guard that the test above has teeth: a single-pass TF32 product misses the budget by far."""
    M = K = N = 512
    a, b, _, _ = operands(B, M, K, N, "NN", seed=3)
    trunc = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    crude = trunc(a).astype(np.float64) @ trunc(b).astype(np.float64)
    truth = a.astype(np.float64) @ b.astype(np.float64)
    assert np.abs(crude - truth).max() > 50 * 1e-5 * np.sqrt(K)


@pytest.mark.parametrize("layout", ["NN", "NT", "TN"])
def test_accumulate_into_c(B, force_tc, layout):
    from minidiff_b200.backend import functions as F

    M, K, N = 300, 260, 270
    a, b, da, db = operands(B, M, K, N, layout, seed=11)
    c0 = np.random.default_rng(5).standard_normal((M, N)).astype(np.float32)
    dc = B.asarray(c0.copy())
    F._gemm(da, db, out=dc, accumulate=True)
    np.testing.assert_allclose(dc.numpy(), c0 + a @ b, rtol=1e-4, atol=1e-5 * np.sqrt(K))


def test_special_values_and_zero_padding(B, force_tc):
    M, K, N = 160, 136, 136
    a, b, _, _ = operands(B, M, K, N, "NN", seed=2)
    a[3, 5] = np.inf
    b[7, 9] = np.nan
    got = B.matmul(B.asarray(a), B.asarray(b)).numpy()
    want = a @ b
    # NaN inputs poison exactly the entries IEEE says; every entry untouched by a special value is
    # a normal 3xTF32 result.  An inf input keeps its row non-finite, but the split arithmetic may
    # turn IEEE's +-inf into NaN there (inf*b_hi + inf*b_lo can be inf - inf): documented limit.
    assert np.isnan(got[:, 9]).all() and np.isnan(want[:, 9]).all()
    assert not np.isfinite(got[3]).any() and not np.isfinite(want[3]).any()
    ok = np.isfinite(want)
    assert np.isfinite(got[ok]).all()
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-4, atol=1e-4)


def test_small_and_odd_problems_take_the_cuda_core_path(B):
    for M, K, N in [(10, 30, 20), (3, 5, 7), (33, 31, 35), (1, 64, 1)]:
        a, b, da, db = operands(B, M, K, N, "NN", seed=1)
        np.testing.assert_allclose(B.matmul(da, db).numpy(), a @ b, rtol=1e-5, atol=1e-5)
    # strided (non-unit on both axes) operands are gathered by the split pre-pass
    big = np.random.default_rng(0).standard_normal((512, 512)).astype(np.float32)
    d = B.asarray(big)
    got = B.matmul(d[::2, ::2], d[1::2, ::2].T).numpy()
    np.testing.assert_allclose(got, big[::2, ::2] @ big[1::2, ::2].T, rtol=1e-4, atol=2e-4)


def test_launch_is_tensor_core_for_large_shapes(B):
    """The auto dispatcher must pick tcgen05 for BASELINE-like shapes (no silent SIMT fallback)."""
    from minidiff_b200.backend._lib import check, lib

    a, b, da, db = operands(B, 1024, 1024, 1024, "NN", seed=9)
    B.matmul(da, db)                      # first launch pays module load + attribute setup
    check(lib.mdb_prof_enable(1))
    B.matmul(da, db)
    ms, n, fl = C.c_double(), C.c_uint64(), C.c_double()
    check(lib.mdb_prof_read(2, C.byref(ms), C.byref(n), C.byref(fl)))
    check(lib.mdb_prof_enable(0))
    assert n.value == 1 and fl.value == 2.0 * 1024 ** 3
    assert fl.value / (ms.value * 1e-3) > 40e12, f"{fl.value / ms.value / 1e9:.1f} TFLOP/s: not the tensor-core path"


@pytest.mark.parametrize("layout", ["NN", "NT", "TN", "TT"])
@pytest.mark.parametrize("variant", [131072, 2097152])
def test_experimental_pair_variants_are_correct(B, layout, variant):
    """The A-operand-in-tensor-memory kernel (mdb_gemm_tune bit 17) and the 4-CTA-cluster kernel with
    TMA-multicast A tiles (bit 21) are never selected by the dispatcher (both measured slower than the
    plain pair kernel), but they must stay correct references."""
    from minidiff_b200.backend._lib import check, lib

    check(lib.mdb_gemm_config(2))
    check(lib.mdb_gemm_tune(4 | 32 | variant))
    try:
        for M, K, N in [(300, 260, 272), (1024, 1024, 768), (520, 96, 1030)]:
            a, b, da, db = operands(B, M, K, N, layout, seed=M + N)
            np.testing.assert_allclose(B.matmul(da, db).numpy(), a @ b, rtol=1e-4, atol=1e-5 * np.sqrt(K))
    finally:
        check(lib.mdb_gemm_tune(4))
        check(lib.mdb_gemm_config(0))


def test_random_shapes_fuzz(B):
    """Seeded fuzz over the dispatcher (pair / single / CUDA-core kernels chosen automatically):
    random extents (ragged against every tile size), random operand layouts, plain and accumulate."""
    from minidiff_b200.backend import functions as F

    rng = np.random.default_rng(2026)
    for case in range(36):
        M, N = int(rng.integers(33, 1400)), int(rng.integers(33, 1400))
        K = int(rng.integers(32, 2600))
        layout = ["NN", "NT", "TN", "TT"][int(rng.integers(4))]
        a, b, da, db = operands(B, M, K, N, layout, seed=1000 + case)
        truth = a.astype(np.float64) @ b.astype(np.float64)
        tol = dict(rtol=1e-4, atol=1.5e-5 * np.sqrt(K))
        if case % 3 == 2:
            c0 = rng.standard_normal((M, N)).astype(np.float32)
            dc = B.asarray(c0.copy())
            F._gemm(da, db, out=dc, accumulate=True)
            np.testing.assert_allclose(dc.numpy(), c0 + truth, err_msg=f"{M}x{K}x{N} {layout} acc", **tol)
        else:
            np.testing.assert_allclose(B.matmul(da, db).numpy(), truth, err_msg=f"{M}x{K}x{N} {layout}", **tol)
