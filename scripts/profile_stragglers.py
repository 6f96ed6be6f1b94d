"""The kernels below 0.9 of the HBM peak, a few launches each, for `ncu --set full`:
    python scripts/profile_stragglers.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import minidiff_b200.backend as B

rng = np.random.default_rng(0)
t = B.asarray(rng.standard_normal((8192, 8192), dtype=np.float32))
h = B.asarray(rng.standard_normal((16384, 4096), dtype=np.float32))
for _ in range(2):
    B.sum(t, axis=0); B.sum(h, axis=0); B.exp(t); B.sum(t); B.sum(t, axis=1); B.sin(t)
B.synchronize()
print("done")
