"""Small CTA-pair GEMMs (data-parallel, stream-K, fused epilogue) for compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/racecheck_gemm.py
    compute-sanitizer --tool memcheck  python scripts/racecheck_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend import functions as F
from minidiff_b200.backend._lib import check, lib

check(lib.mdb_gemm_config(2)); check(lib.mdb_gemm_tune(4 | 32))
rng = np.random.default_rng(0)
for (M, K, N), sk in (((300, 160, 272), 0), ((520, 640, 1030 + 2), 1), ((1300, 512, 520), 1)):
    check(lib.mdb_gemm_knob(5, sk))
    a, b = rng.standard_normal((M, K), dtype=np.float32), rng.standard_normal((K, N), dtype=np.float32)
    got = B.matmul(B.asarray(a), B.asarray(b)).numpy()
    np.testing.assert_allclose(got, a.astype(np.float64) @ b, rtol=1e-4, atol=1e-5 * np.sqrt(K))
    bias = B.asarray(rng.standard_normal(N).astype(np.float32))
    out = F._gemm_fused(B.asarray(a), B.asarray(b), bias=bias, relu=True)
    assert out is not None and np.isfinite(out.numpy()).all()
print("racecheck workload OK")
