"""Per-op HBM efficiency of the elementwise / reduction kernels at BASELINE sizes (algorithmic
bytes / CUDA-event time via the library profiler).   python scripts/microbench_ew.py [reps]"""
import ctypes as C
import json
import sys
sys.path.insert(0, ".")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
from minidiff_b200.backend import functions as F

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6549.8
N = M = 8192
rng = np.random.default_rng(0)
t = B.asarray(rng.standard_normal((N, M), dtype=np.float32))
u = B.asarray(rng.standard_normal((N, M), dtype=np.float32))
a = B.asarray(rng.standard_normal((N, 1), dtype=np.float32))
c = B.asarray(rng.standard_normal((1, M), dtype=np.float32))
mask = B.greater(t, 0)
g0 = B.broadcast_to(B.asarray(np.float32(1.0)), (N, M))
out_a = B.zeros((N, 1), dtype=np.float32)
out_c = B.zeros((1, M), dtype=np.float32)
h = B.asarray(rng.standard_normal((16384, 4096), dtype=np.float32))
bias = B.asarray(rng.standard_normal((4096,), dtype=np.float32))


def ered(op, out, *ins, acc=0):
    n = len(ins)
    descs = (F.MdbArray * n)()
    for i, o in enumerate(ins):
        if isinstance(o, B.DeviceArray):
            descs[i] = o.d
        else:
            F._fill_imm(descs[i], o)
    check(lib.mdb_elementwise_reduce(F.OP[op], C.byref(out.d), n, descs, acc))


CASES = [
    ("outer a*c            (w E)", 0, lambda: B.multiply(a, c)),
    ("t + a colvec      (r E w E)", 0, lambda: B.add(t, a)),
    ("t + c rowvec      (r E w E)", 0, lambda: B.add(t, c)),
    ("t + u             (r2E w E)", 0, lambda: B.add(t, u)),
    ("sin(t)            (r E w E)", 0, lambda: B.sin(t)),
    ("exp(t)            (r E w E)", 0, lambda: B.exp(t)),
    ("t**2              (r E w E)", 0, lambda: B.power(t, 2)),
    ("2.0*t             (r E w E)", 0, lambda: B.multiply(2.0, t)),
    ("t > 0         (r E w E/4)", 0, lambda: B.greater(t, 0)),
    ("where(m,t,0) (r 1.25E w E)", 0, lambda: B.where(mask, t, 0)),
    ("t * mask     (r 1.25E w E)", 0, lambda: B.multiply(t, mask)),
    ("t += u in place   (r2E w E)", 0, lambda: t.__iadd__(u)),
    ("POW_BWD(g0,t,2)   (r E w E)", 0, lambda: F._launch_ew("POW_BWD", B.DeviceArray.empty((N, M), np.float32), [g0, t, 2])),
    ("SIN_BWD(u,t)      (r2E w E)", 0, lambda: F._launch_ew("SIN_BWD", B.DeviceArray.empty((N, M), np.float32), [u, t])),
    ("h + bias (16384x4096)", 0, lambda: B.add(h, bias)),
    ("ones_like fill        (w E)", 0, lambda: B.ones_like(t)),
    ("sum(t) full           (r E)", 1, lambda: B.sum(t)),
    ("sum(t, axis=1)        (r E)", 1, lambda: B.sum(t, axis=1)),
    ("sum(t, axis=0)        (r E)", 1, lambda: B.sum(t, axis=0)),
    ("sum(t*c, axis=1) fused(r E)", 1, lambda: ered("MUL", out_a, t, c)),
    ("sum(t*a, axis=0) fused(r E)", 1, lambda: ered("MUL", out_c, t, a)),
    ("sum(h, axis=0) bias grad", 1, lambda: B.sum(h, axis=0)),
    ("mean(t)               (r E)", 1, lambda: B.mean(t)),
]
print(f"{'case':32s} {'us':>9s} {'GB/s':>9s} {'of peak':>8s}")
for name, cls, fn in CASES:
    for _ in range(2):
        fn()
    check(lib.mdb_prof_enable(1))
    for _ in range(reps):
        fn()
    ms, n, w = C.c_double(), C.c_uint64(), C.c_double()
    check(lib.mdb_prof_read(cls, C.byref(ms), C.byref(n), C.byref(w)))
    check(lib.mdb_prof_enable(0))
    gbs = w.value / (ms.value * 1e-3) / 1e9
    print(f"{name:32s} {ms.value / n.value * 1e3:9.1f} {gbs:9.0f} {gbs / PEAK:8.2f}")
