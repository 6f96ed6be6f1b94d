"""Per-role stall counters of the CTA-pair kernel (MDB_GEMM_TIMING=1 instantiation) for kernel-flag variants.
    MDB_GEMM_TIMING=1 python scripts/gemm_stalls.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend import functions as F
from minidiff_b200.backend._lib import check, lib

variants = [("base", 4 | 32)] + [(n, 4 | 32 | int(f)) for n, f in (a.split("=") for a in sys.argv[1:])]
for M, K, N in ((16384, 4096, 4096), (16384, 1024, 4096)):
    a = B.asarray(np.random.default_rng(0).standard_normal((M, K), dtype=np.float32))
    b = B.asarray(np.random.default_rng(1).standard_normal((K, N), dtype=np.float32))
    out = B.zeros((M, N), dtype=np.float32)
    for name, fl in variants:
        check(lib.mdb_gemm_tune(fl))
        for _ in range(2):
            print(f"--- {name} {M}x{K}x{N}", file=sys.stderr, flush=True)
            F._gemm(a, b, out=out)
        B.synchronize()
