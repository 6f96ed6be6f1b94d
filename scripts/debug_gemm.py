"""Probe the tcgen05 GEMM with structured operands (debug aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check

check(lib.mdb_gemm_config(2))
np.set_printoptions(linewidth=200, precision=3, suppress=True)


def run(a, b, layout):
    da = B.asarray(a) if layout[0] == "N" else B.asarray(np.ascontiguousarray(a.T)).T
    db = B.asarray(b) if layout[1] == "N" else B.asarray(np.ascontiguousarray(b.T)).T
    return B.matmul(da, db).numpy()


M = K = N = 128
rng = np.random.default_rng(0)
for layout in ("NT", "NN", "TN", "TT"):
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    got = run(a, b, layout)
    want = a @ b
    print(f"== {layout}: max|got|={np.abs(got).max():.4g} max|want|={np.abs(want).max():.4g} "
          f"max err={np.abs(got-want).max():.4g} zeros={np.mean(got == 0):.3f} nan={np.isnan(got).mean():.3f}")
    if np.abs(got - want).max() > 1e-3:
        # which rows / cols are right?
        good = np.abs(got - want) < 1e-3
        print("   good fraction", good.mean(), "good rows", np.where(good.all(axis=1))[0][:10],
              "good cols", np.where(good.all(axis=0))[0][:10])
        print("   got[:4,:6]\n", got[:4, :6], "\n   want[:4,:6]\n", want[:4, :6])
        # identity probes: A = I  -> C = B ; B = I -> C = A
        eye = np.eye(128, dtype=np.float32)
        pat = (np.arange(128)[:, None] * 1000 + np.arange(128)[None, :]).astype(np.float32)
        g1 = run(eye, pat, layout)
        print("   A=I, B[k,n]=1000k+n  -> got[:3,:6]\n", g1[:3, :6], "\n   rows 8,32,64:\n", g1[[8, 32, 64], :6])
        g2 = run(pat, eye, layout)
        print("   A[m,k]=1000m+k, B=I  -> got[:3,:6]\n", g2[:3, :6], "\n   rows 8,32,64:\n", g2[[8, 32, 64], :6])
