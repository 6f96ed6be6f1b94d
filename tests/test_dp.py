"""Data-parallel exchange (BASELINE config 4).

* CPU (`-m "not gpu"`): world-size-2 gloo run of the gradient-averaging maths on oracle gradients.
* GPU: 2 ranks / 2 GPUs through NCCL vs one GPU on the full batch (skipped with < 2 GPUs)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "dp_worker.py")


def launch(mode, out, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode, str(out)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300)


def test_gradient_averaging_over_two_gloo_ranks(tmp_path):
    out = tmp_path / "ok.txt"
    r = launch("gloo-oracle", out, 2, 29631)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert out.read_text() == "OK"


@pytest.mark.gpu
def test_two_gpu_step_matches_single_gpu(tmp_path):
    import np_minidiff as orc

    import minidiff_b200 as md
    from minidiff_b200 import workloads as W
    from minidiff_b200.backend._lib import lib

    n = C.c_int(0)
    lib.mdb_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "dp.npz"
    r = launch("nccl", out, 2, 29633)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    got = np.load(out)
    dims, batch = (32, 64, 64, 16), 256
    X, Y = orc.mlp_data(batch, dims[0], dims[-1])
    params = [md.Tensor(p.copy(), allow_grad=True) for p in orc.mlp_params(dims)]
    for _ in range(2):
        W.mlp_train_step(md.Tensor(X), md.Tensor(Y), params, 0.01)
    for i, p in enumerate(params):
        np.testing.assert_allclose(got[f"arr_{i}"], p.as_numpy(), rtol=1e-4, atol=1e-6)
    # aliased-gradient scenario of the worker: expected averages computed with NumPy per shard
    rng = np.random.default_rng(77)
    Wa = rng.standard_normal((dims[0], 1024)).astype(np.float32)
    Pa = rng.standard_normal((batch // 2, 1024)).astype(np.float32)
    dW, dP = 0, 0
    for r_ in range(2):
        Xs = X[r_ * batch // 2:(r_ + 1) * batch // 2].astype(np.float64)
        h = Xs @ Wa + Pa
        g = 2 * h / h.size
        dW, dP = dW + Xs.T @ g / 2, dP + g / 2
    np.testing.assert_allclose(got["alias_dW"], dW, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["alias_dP"], dP, rtol=1e-4, atol=1e-7)
