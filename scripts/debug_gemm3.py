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
mode = os.environ.get("MDB_GEMM_DEBUG", "0")
got = B.matmul(B.asarray(a), B.asarray(b)).numpy()     # NN: A K-major, B MN-major
print("NN mode", mode, "max", np.abs(got).max(), "zeros", (got == 0).mean())
if mode == "2":
    print("B_hi smem dump rows 0..3 (chunk 0, k rows 0..3 of k-block 1):\n", got[:4, 32:64])
    print("rows 8, 31, 32, 33, 64, 96:\n", got[[8, 31, 32, 33, 64, 96], 32:64])
else:
    want = a.astype(np.float64) @ b.astype(np.float64)
    print("got[:3,:6]\n", got[:3, :6], "\nwant[:3,:6]\n", want[:3, :6])
    print("relerr max", np.abs(got - want).max() / np.abs(want).max())
