"""A/B timing of the CTA-pair GEMM planner knobs (tile order, L2 hints, stream-K; SWEEP=flags: kernel flags;
SWEEP=split: operand split x chain length) on the GEMM shapes of the C4 step (per-GPU batch B) and C3.
    python scripts/gemm_sweep.py [B] [rounds] [shape,shape..]
All variants of a shape are timed INTERLEAVED (round-robin, several rounds, after a warm-up that
brings the chip to its power-capped steady state), so they see the same clocks: a sequential sweep
gave the first variant 10-15 % for free.  Prints the median ms and TFLOP/s (2MNK) per variant."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minidiff_b200.backend as Bk  # noqa: E402
from minidiff_b200.backend import functions as F  # noqa: E402
from minidiff_b200.backend._lib import check, lib  # noqa: E402

KNOB = dict(raster=0, group=1, hint_a=2, hint_b=3, hint_c=4, streamk=5, l2=6, split=8, chunk=9)
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 7
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None


def setk(**kw):
    for k in KNOB.values():
        check(lib.mdb_gemm_knob(k, -1))
    for k, v in kw.items():
        check(lib.mdb_gemm_knob(KNOB[k], v))


def ev():
    e = C.c_void_p()
    check(lib.mdb_event_create(C.byref(e)))
    return e


E0, E1 = ev(), ev()


def timeit(fn, reps):
    check(lib.mdb_event_record(E0))
    for _ in range(reps):
        fn()
    check(lib.mdb_event_record(E1))
    ms = C.c_float()
    check(lib.mdb_event_elapsed_ms(E0, E1, C.byref(ms)))
    return ms.value / reps


def rnd(shape):
    return Bk.asarray(np.random.default_rng(sum(shape)).standard_normal(shape, dtype=np.float32))


def plan():
    out = (C.c_int * 8)()
    lib.mdb_gemm_last_plan(out)
    return f"cl{out[0]} r{out[1]}g{out[2]} dp{out[3]} sk{out[4]}x{out[5]} h{out[6]:03d}"


shapes = [("fwd1", (Bsz, 1024, 4096), 0, 0), ("fwd2", (Bsz, 4096, 4096), 0, 0), ("fwd3", (Bsz, 4096, 1024), 0, 0),
          ("dW3", (4096, Bsz, 1024), 1, 0), ("dh2", (Bsz, 1024, 4096), 0, 1), ("dW2", (4096, Bsz, 4096), 1, 0),
          ("dh1", (Bsz, 4096, 4096), 0, 1), ("dW1", (1024, Bsz, 4096), 1, 0), ("c3", (8192, 8192, 8192), 0, 0)]
variants = [("old_r0g8", dict(raster=0, group=8, hint_a=0, hint_b=0, hint_c=0, streamk=0)),
            ("auto", dict()),
            ("auto_nosk", dict(streamk=0)),
            ("auto_nohint", dict(hint_a=0, hint_b=0, hint_c=0)),
            ("old+hintC", dict(raster=0, group=8, hint_a=0, hint_b=0, hint_c=1, streamk=0)),
            ("old+sk", dict(raster=0, group=8, hint_a=0, hint_b=0, hint_c=0)),
            ("r1g4", dict(raster=1, group=4, streamk=0)),
            ("r1g8", dict(raster=1, group=8, streamk=0)),
            ("r1g16", dict(raster=1, group=16, streamk=0)),
            ("r0g4", dict(raster=0, group=4, streamk=0)),
            ("r0g16", dict(raster=0, group=16, streamk=0))]
if os.environ.get("SWEEP") == "flags":
    # kernel-flag A/B (mdb_gemm_tune bits) instead of planner knobs
    variants = [("base", dict(_flags=4 | 32)), ("lo_rounded", dict(_flags=32)), ("base2", dict(_flags=4 | 32))]
if os.environ.get("SWEEP") == "split":
    # operand split A/B: 3xTF32 vs TF32 + 2 BF16 cross terms
    variants = [("3xtf32 c4", dict(split=0, chunk=4)), ("3xtf32 c2", dict(split=0, chunk=2)), ("hybrid c4", dict(split=1, chunk=4)),
                ("hybrid c2", dict(split=1, chunk=2)), ("hybrid c1", dict(split=1, chunk=1))]
check(lib.mdb_gemm_tune(4 | 32))
for name, (M, K, N), ta, tb in shapes:
    if only and name not in only:
        continue
    a = rnd((K, M)).T if ta else rnd((M, K))
    b = rnd((N, K)).T if tb else rnd((K, N))
    out = Bk.zeros((M, N), dtype=np.float32)
    fn = lambda: F._gemm(a, b, out=out)
    flops = 2.0 * M * K * N
    reps = max(3, int(25e-3 / (flops / 270e12)))          # ~25 ms per measurement
    setk()
    for _ in range(int(1.5 / (flops / 270e12)) + 1):       # ~1.5 s warm-up: steady state under the power cap
        fn()
    res = {vn: [] for vn, _ in variants}
    plans = {}
    for _ in range(rounds):
        for vn, kw in variants:
            kw = dict(kw)
            check(lib.mdb_gemm_tune(kw.pop("_flags", 4 | 32)))
            setk(**kw)
            res[vn].append(timeit(fn, reps))
            plans[vn] = plan()
    print(f"{name:5s} {M}x{K}x{N}  (median of {rounds} interleaved rounds x {reps} launches)")
    for vn, _ in variants:
        ms = float(np.median(res[vn]))
        print(f"    {vn:12s} {ms:8.3f} ms  {flops / ms / 1e9:6.1f} TF  [{plans[vn]}]", flush=True)
    del a, b, out
setk()
