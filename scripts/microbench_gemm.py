"""A/B timing of GEMM kernel variants in ONE process (boxes differ by ~10 % under the power cap).
    python scripts/microbench_gemm.py [n] [reps]"""
import ctypes as C, sys
sys.path.insert(0, ".")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
rng = np.random.default_rng(0)
a = B.asarray(rng.standard_normal((n, n), dtype=np.float32))
b = B.asarray(rng.standard_normal((n, n), dtype=np.float32))
layouts = {"NN": (a, b), "NT": (a, b.T), "TN": (a.T, b)}
def timeit(x, y):
    B.matmul(x, y)
    check(lib.mdb_prof_enable(1))
    for _ in range(reps):
        B.matmul(x, y)
    ms, cnt, fl = C.c_double(), C.c_uint64(), C.c_double()
    check(lib.mdb_prof_read(2, C.byref(ms), C.byref(cnt), C.byref(fl)))
    check(lib.mdb_prof_enable(0))
    return ms.value / cnt.value
for rnd in range(2):
    for flags in (0, 4, 2):
        check(lib.mdb_gemm_tune(flags))
        t = {k: timeit(*v) for k, v in layouts.items()}
        tot = sum(t.values())
        got = B.matmul(a[:1024], b[:, :1024]).numpy()
        truth = a[:1024].numpy().astype(np.float64) @ b[:, :1024].numpy().astype(np.float64)
        err = np.abs(got - truth).max()
        print(f"round {rnd} flags={flags} (cvt_rna={(flags>>1)&1} no_lo_round={(flags>>2)&1}) maxerr {err:.2e}: " +
              " ".join(f"{k} {v:.3f} ms" for k, v in t.items()) + f" | sum {tot:.2f} ms = {3*2*n**3/tot/1e9:.1f} TF/s", flush=True)
