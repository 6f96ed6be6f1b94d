"""Host-only microbenchmarks of single ops against the null device (see host_profile.py)."""
import sys, time
sys.path.insert(0, "scripts")
import host_profile
sys.argv = sys.argv[:1]
host_profile.prepare()
import numpy as np
import minidiff_b200 as md
from minidiff_b200.backend import functions as F
from minidiff_b200.backend.device_array import DeviceArray, F32


def bench(name, fn, n=100000):
    for _ in range(2000):
        fn()
    best = 1e9
    for _ in range(5):                       # best of 5 batches: the container's load varies
        t = time.perf_counter()
        for _ in range(n // 5):
            fn()
        best = min(best, (time.perf_counter() - t) / (n // 5))
    print(f"{name:34s} {best * 1e6:6.2f} us")


a = DeviceArray.empty((2, 4), F32); b = DeviceArray.empty((2, 4), F32)
x = md.Tensor(a, allow_grad=True); y = md.Tensor(b, allow_grad=True)
bench("DeviceArray.empty", lambda: DeviceArray.empty((2, 4), F32))
bench("F.add(a, b)", lambda: F.add(a, b))
bench("F.multiply(a, 2)", lambda: F.multiply(a, 2))
bench("F.sin(a)", lambda: F.sin(a))
bench("md.Tensor(a)", lambda: md.Tensor(a))
bench("md.add(x, y)", lambda: md.add(x, y))
bench("x * y", lambda: x * y)
bench("2 * y", lambda: 2 * y)
bench("md.sin(x)", lambda: md.sin(x))
bench("x ** 2", lambda: x ** 2)
with md.no_grad():
    bench("x * y (no_grad)", lambda: x * y)


def c1():
    f = 2 * y * md.sin(x) - x ** 2
    f.backward(allow_higher_order=True)
    x.grad.backward()


def fwd():
    return 2 * y * md.sin(x) - x ** 2


def fwd_bwd1():
    fwd().backward()


bench("C1 forward (5 ops)", fwd, 20000)
bench("C1 fwd + 1st-order bwd", fwd_bwd1, 20000)
bench("C1 full (29 launches)", c1, 5000)

if "--profile" in sys.argv or True:
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(1500):
        c1()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
