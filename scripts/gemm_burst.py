"""Burst (not power-capped) timing of the CTA-pair GEMM: one launch after 1 s of idle, repeated; the
sustained figure comes from scripts/gemm_sweep.py.     python scripts/gemm_burst.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend import functions as F
from minidiff_b200.backend._lib import check, lib


def ev():
    e = C.c_void_p(); check(lib.mdb_event_create(C.byref(e))); return e


e0, e1 = ev(), ev()
for (M, K, N) in ((8192, 8192, 8192), (65536, 4096, 4096), (65536, 1024, 4096), (4096, 65536, 4096)):
    rng = np.random.default_rng(0)
    a = B.asarray(rng.standard_normal((K, M), dtype=np.float32)).T if K > M else B.asarray(rng.standard_normal((M, K), dtype=np.float32))
    b = B.asarray(rng.standard_normal((K, N), dtype=np.float32))
    out = B.zeros((M, N), dtype=np.float32)
    F._gemm(a, b, out=out); B.synchronize()
    ts = []
    for _ in range(5):
        time.sleep(1.0)
        check(lib.mdb_event_record(e0)); F._gemm(a, b, out=out); check(lib.mdb_event_record(e1))
        ms = C.c_float(); check(lib.mdb_event_elapsed_ms(e0, e1, C.byref(ms))); ts.append(ms.value)
    best = min(ts)
    print(f"{M}x{K}x{N}: burst best {best:.3f} ms = {2.0*M*K*N/best/1e9:.1f} TFLOP/s fp32-equivalent ({3*2.0*M*K*N/best/1e9:.0f} on the tensor pipe); all {[round(t,3) for t in ts]}", flush=True)
    del a, b, out
