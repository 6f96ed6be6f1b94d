import os, sys
import numpy as np
sys.path.insert(0, ".")
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
check(lib.mdb_gemm_config(2))
np.set_printoptions(linewidth=220, precision=2, suppress=True)
M = K = N = 128
a = (np.arange(M)[:, None] * 1000 + np.arange(K)[None, :]).astype(np.float32)      # A[m,k] = 1000m + k
b = (np.arange(K)[:, None] * 1000 + np.arange(N)[None, :]).astype(np.float32)      # B[k,n] = 1000k + n
bt = np.ascontiguousarray(b.T)
mode = os.environ.get("MDB_GEMM_DEBUG", "0")
got = B.matmul(B.asarray(a), B.asarray(bt).T).numpy()     # NT: both K-major
print("mode", mode, "max", np.abs(got).max(), "zeros", (got == 0).mean())
print("rows 0..3, cols 0..8:\n", got[:4, :8])
print("rows 0..3, cols 32..40:\n", got[:4, 32:40])
print("rows 8,9 cols 0..8:\n", got[8:10, :8], "\nrow 1 cols 0..32:\n", got[1, :32])
print("cols 64..72:\n", got[:2, 64:72])
