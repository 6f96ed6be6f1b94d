"""CTA-pair (cta_group::2) GEMM: correctness against float64 and A/B timing against the single-CTA
kernel in ONE process.   python scripts/debug_pair.py [quick]"""
import ctypes as C, sys
sys.path.insert(0, ".")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
skip_check = len(sys.argv) > 2
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
from minidiff_b200.backend import functions as F

check(lib.mdb_gemm_config(2))
rng = np.random.default_rng(0)


def ops(M, K, N, layout):
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    da = B.asarray(a) if layout[0] == "N" else B.asarray(np.ascontiguousarray(a.T)).T
    db = B.asarray(b) if layout[1] == "N" else B.asarray(np.ascontiguousarray(b.T)).T
    return a, b, da, db


bad = 0
for (M, K, N) in [] if skip_check else [(256, 128, 256), (256, 512, 512), (512, 96, 768), (300, 260, 272), (1000, 776, 556),
                  (2048, 1024, 1536), (129, 200, 132), (4096, 4096, 1024)]:
    for layout in ("NN", "NT", "TN", "TT"):
        a, b, da, db = ops(M, K, N, layout)
        truth = a.astype(np.float64) @ b.astype(np.float64)
        res = {}
        for name, fl in (("pair", 4 | 32), ("single", 4 | 16), ("mc", 4 | 32 | 2097152)):
            check(lib.mdb_gemm_tune(fl))
            res[name] = np.abs(B.matmul(da, db).numpy() - truth).max()
        check(lib.mdb_gemm_tune(4 | 32))
        c0 = rng.standard_normal((M, N)).astype(np.float32)
        dc = B.asarray(c0.copy())
        F._gemm(da, db, out=dc, accumulate=True)
        acc_err = np.abs(dc.numpy() - (c0 + truth)).max()
        ok = res["pair"] <= 2 * res["single"] + 1e-6 and acc_err <= 2 * res["single"] + 1e-5 and res["mc"] <= 2 * res["single"] + 1e-6
        bad += (not ok)
        print(f"{M}x{K}x{N} {layout}: mc {res['mc']:.2e} pair {res['pair']:.2e} single {res['single']:.2e} acc {acc_err:.2e} {'ok' if ok else 'BAD'}", flush=True)
print("correctness:", "ALL OK" if not bad else f"{bad} BAD", flush=True)
if bad:
    sys.exit(1)


def timeit(x, y, reps):
    B.matmul(x, y)
    check(lib.mdb_prof_enable(1))
    for _ in range(reps):
        B.matmul(x, y)
    ms, cnt, fl = C.c_double(), C.c_uint64(), C.c_double()
    check(lib.mdb_prof_read(2, C.byref(ms), C.byref(cnt), C.byref(fl)))
    check(lib.mdb_prof_enable(0))
    return ms.value / cnt.value


cases = [("8192^3 NN", 8192, 8192, 8192, "NN"), ("8192^3 NT", 8192, 8192, 8192, "NT"), ("8192^3 TN", 8192, 8192, 8192, "TN")]
if not quick:
    cases += [("fwd1 65536x1024x4096 NN", 65536, 1024, 4096, "NN"), ("fwd2 65536x4096x4096 NN", 65536, 4096, 4096, "NN"),
              ("fwd3 65536x4096x1024 NN", 65536, 4096, 1024, "NN"), ("dh2 65536x1024x4096 NT", 65536, 1024, 4096, "NT"),
              ("dW1 1024x65536x4096 TN", 1024, 65536, 4096, "TN"), ("dW2 4096x65536x4096 TN", 4096, 65536, 4096, "TN"),
              ("dW3 4096x65536x1024 TN", 4096, 65536, 1024, "TN"),
              ("fwd2/8 8192x4096x4096 NN", 8192, 4096, 4096, "NN"), ("dW2/8 4096x8192x4096 TN", 4096, 8192, 4096, "TN"),
              ("fwd3/8 8192x4096x1024 NN", 8192, 4096, 1024, "NN")]
for name, M, K, N, layout in cases:
    a = B.asarray(rng.standard_normal((M, K), dtype=np.float32)) if layout[0] == "N" else B.asarray(rng.standard_normal((K, M), dtype=np.float32)).T
    b = B.asarray(rng.standard_normal((K, N), dtype=np.float32)) if layout[1] == "N" else B.asarray(rng.standard_normal((N, K), dtype=np.float32)).T
    out = []
    for rnd in range(1):
        variants = (("pair", 4 | 32), ("mc", 4 | 32 | 2097152), ("single", 4 | 16))
        if skip_check:
            variants = (("pair", 4 | 32), ("mc", 4 | 32 | 2097152))
        for nm, fl in variants:
            check(lib.mdb_gemm_tune(fl))
            t = timeit(a, b, 5)
            out.append(f"{nm} {t:.3f} ms {2.0*M*N*K/t/1e9:.0f} TF/s")
    print(name, "|", " | ".join(out), flush=True)
    del a, b
