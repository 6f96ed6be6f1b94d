"""Column sums (bias gradients) at the C4 shapes: python scripts/microbench_colsum.py"""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check

def ev():
    e = C.c_void_p(); check(lib.mdb_event_create(C.byref(e))); return e
E0, E1 = ev(), ev()
for shape in ((65536, 4096), (65536, 1024), (16384, 4096), (8192, 4096), (8192, 1024), (8192, 8192), (4096, 512)):
    x = B.asarray(np.random.default_rng(0).standard_normal(shape, dtype=np.float32))
    for _ in range(3): B.sum(x, axis=0)
    check(lib.mdb_event_record(E0))
    for _ in range(20): B.sum(x, axis=0)
    check(lib.mdb_event_record(E1))
    ms = C.c_float(); check(lib.mdb_event_elapsed_ms(E0, E1, C.byref(ms)))
    us = ms.value / 20 * 1e3
    print(f"{shape}: {us:8.1f} us  {x.nbytes / us / 1e3:7.0f} GB/s  {x.nbytes / us / 1e3 / 6549.8:.2f}", flush=True)
    del x
