"""Per-role stall breakdown of the CTA-pair GEMM (MDB_GEMM_TIMING=1 diagnostic build of the kernel).
    MDB_GEMM_TIMING=1 python scripts/gemm_timing.py [n]"""
import sys
sys.path.insert(0, ".")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
check(lib.mdb_gemm_config(2))
rng = np.random.default_rng(0)
a = B.asarray(rng.standard_normal((n, n), dtype=np.float32))
b = B.asarray(rng.standard_normal((n, n), dtype=np.float32))
nc = 4 | 32 | 512 | 256
# 512 = no conversion, 256 = hi*hi MMAs only, 128 = alternate accumulators, 4096 = L2-resident k range,
# 16384 = .release.cluster remote arrives (all wrong-result switches exist in the diagnostic build only)
for name, fl in (("pair_ts", 4 | 32 | 131072), ("pair_ts", 4 | 32 | 131072), ("pair43", 4 | 32), ("pair43", 4 | 32), ("pair52", 4 | 32 | 64), ("noconv43", 4 | 32 | 512), ("hihi43", 4 | 32 | 256),
                 ("nc-hihi43", nc), ("release-arrive", 4 | 32 | 16384), ("A-from-TMEM", 4 | 32 | 65536),
                 ("A-from-TMEM-noconv", 4 | 32 | 65536 | 512)):
    check(lib.mdb_gemm_tune(fl))
    print("==", name, file=sys.stderr, flush=True)
    B.matmul(a, b)
    B.synchronize()
