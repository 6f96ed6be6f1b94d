"""DRAM traffic of the CTA-pair GEMM per planner variant, for ncu:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
        -k regex:gemm_3xtf32_pair --csv --log-file out.csv python scripts/gemm_traffic.py [B]
Each (shape, variant) is launched exactly once after the buffers exist; the launch order is printed
so the CSV rows can be labelled (scripts/ncu_traffic_table.py)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minidiff_b200.backend as Bk  # noqa: E402
from minidiff_b200.backend import functions as F  # noqa: E402
from minidiff_b200.backend._lib import check, lib  # noqa: E402

KNOB = dict(raster=0, group=1, hint_a=2, hint_b=3, hint_c=4, streamk=5, l2=6)
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536


def setk(**kw):
    for k in KNOB.values():
        check(lib.mdb_gemm_knob(k, -1))
    for k, v in kw.items():
        check(lib.mdb_gemm_knob(KNOB[k], v))


def rnd(shape):
    return Bk.asarray(np.random.default_rng(sum(shape)).standard_normal(shape, dtype=np.float32))


shapes = [("fwd1", (Bsz, 1024, 4096), 0, 0), ("fwd2", (Bsz, 4096, 4096), 0, 0), ("fwd3", (Bsz, 4096, 1024), 0, 0),
          ("dW3", (4096, Bsz, 1024), 1, 0), ("dh2", (Bsz, 1024, 4096), 0, 1), ("dW2", (4096, Bsz, 4096), 1, 0),
          ("dh1", (Bsz, 4096, 4096), 0, 1), ("dW1", (1024, Bsz, 4096), 1, 0), ("c3", (8192, 8192, 8192), 0, 0)]
variants = [("old_r0g8", dict(raster=0, group=8, hint_a=0, hint_b=0, hint_c=0, streamk=0)),
            ("auto", dict()),
            ("auto_nohint", dict(hint_a=0, hint_b=0, hint_c=0)),
            ("auto_l2_24", dict(l2=24)),
            ("auto_l2_64", dict(l2=64)),
            ("r1g16", dict(raster=1, group=16)),
            ("r1g4", dict(raster=1, group=4)),
            ("r0g4", dict(raster=0, group=4)),
            ("r0g16", dict(raster=0, group=16))]
if len(sys.argv) > 2 and sys.argv[2] == "matrix":
    # raster x group x hints matrix on the K = 4096 shapes (where the panels of one wave exceed the L2)
    shapes = [s_ for s_ in shapes if s_[0] in sys.argv[3].split(",")]
    variants = []
    for r in (1, 0):
        for g in (4, 6, 8, 10, 12, 16):
            for h in ((0, 0, 0), (0, 2, 0), (0, 2, 1), (2, 0, 0), (0, 0, 1)):
                if (r == 1 and h[0] == 2) or (r == 0 and h[1] == 2):
                    continue
                variants.append((f"r{r}g{g}_h{h[0]}{h[1]}{h[2]}", dict(raster=r, group=g, hint_a=h[0], hint_b=h[1], hint_c=h[2], streamk=0)))
check(lib.mdb_gemm_tune(4 | 32))
for name, (M, K, N), ta, tb in shapes:
    a = rnd((K, M)).T if ta else rnd((M, K))
    b = rnd((N, K)).T if tb else rnd((K, N))
    out = Bk.zeros((M, N), dtype=np.float32)
    alg = 4.0 * (M * K + K * N + M * N)
    for vn, kw in variants:
        setk(**kw)
        F._gemm(a, b, out=out)
        p = (C.c_int * 8)()
        lib.mdb_gemm_last_plan(p)
        print(f"LAUNCH {name} {vn} alg_bytes={alg:.0f} plan=cl{p[0]}_r{p[1]}g{p[2]}_dp{p[3]}_sk{p[4]}x{p[5]}_h{p[6]:03d}", flush=True)
    Bk.synchronize()
    del a, b, out
setk()
