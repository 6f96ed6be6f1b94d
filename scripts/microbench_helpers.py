"""Device vs host-NumPy time of the data-dependent helpers / random family / float64 + stacked matmul
 at sizes where the work is bandwidth- or compute-bound.
    python scripts/microbench_helpers.py  -> profiles/r02_microbench_helpers.txt"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B

B.assert_live()
rng = np.random.default_rng(0)


def dev_time(fn, reps=5):
    fn(); B.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    B.synchronize()
    return (time.perf_counter() - t) / reps


def host_time(fn, reps=2):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps


n = 1 << 26
m_np = rng.random(n) < 0.1
v_np = rng.standard_normal(n).astype(np.float32)
m, v = B.asarray(m_np), B.asarray(v_np)
e_np = rng.integers(0, 1 << 20, 1 << 24); t_np = rng.integers(0, 1 << 20, 1024)
e, t = B.asarray(e_np), B.asarray(t_np)
f_np = rng.integers(0, 4096 * 4096, 1 << 24); f = B.asarray(f_np)
w_np = rng.random(4096); w = B.asarray(w_np)
a64, b64 = rng.standard_normal((2048, 2048)), rng.standard_normal((2048, 2048))
da64, db64 = B.asarray(a64), B.asarray(b64)
sa, sb = rng.standard_normal((64, 256, 256)).astype(np.float32), rng.standard_normal((64, 256, 256)).astype(np.float32)
dsa, dsb = B.asarray(sa), B.asarray(sb)
idx_np = rng.integers(0, 1 << 20, 1 << 22); tab_np = rng.standard_normal((1 << 20, 16)).astype(np.float32)
idx, tab = B.asarray(idx_np), B.asarray(tab_np)
cases = [
    ("argwhere, 2^26 bool, 10 % set", lambda: B.argwhere(m), lambda: np.argwhere(m_np)),
    ("v[mask], 2^26 fp32", lambda: v[m], lambda: v_np[m_np]),
    ("isin, 2^24 int64 in 1024", lambda: B.isin(e, t), lambda: np.isin(e_np, t_np)),
    ("unravel_index, 2^24 -> (4096, 4096)", lambda: B.unravel_index(f, (4096, 4096)), lambda: np.unravel_index(f_np, (4096, 4096))),
    ("table[idx] gather, 2^22 rows of 16 fp32 (validated)", lambda: tab[idx], lambda: tab_np[idx_np]),
    ("randint, 2^26 int64", lambda: B.randint(0, 1000, size=n), lambda: np.random.randint(0, 1000, size=n)),
    ("randn, 2^26 float64", lambda: B.randn(n), lambda: np.random.randn(n)),
    ("binomial(20, 0.3), 2^24", lambda: B.binomial(20, 0.3, size=1 << 24), lambda: np.random.binomial(20, 0.3, size=1 << 24)),
    ("permutation, 2^22", lambda: B.permutation(1 << 22), lambda: np.random.permutation(1 << 22)),
    ("choice weighted, 2^22 of 4096", lambda: B.choice(4096, size=1 << 22, p=w), lambda: np.random.choice(4096, size=1 << 22, p=w_np / w_np.sum())),
    ("matmul float64 2048^3", lambda: B.matmul(da64, db64), lambda: a64 @ b64),
    ("matmul stacked 64 x 256^3 fp32 (one launch)", lambda: B.matmul(dsa, dsb), lambda: np.matmul(sa, sb)),
]
print(f"{'case':52s} {'device ms':>10s} {'NumPy ms':>10s} {'ratio':>7s}   (host: {os.cpu_count()} cores)")
for name, d, h in cases:
    td, th = dev_time(d), host_time(h)
    print(f"{name:52s} {td * 1e3:10.3f} {th * 1e3:10.1f} {th / td:7.1f}", flush=True)
