"""Host-side cost per op (Python + ctypes), measured WITHOUT a GPU against a null device.

    python scripts/host_profile.py [c1|c2|mlp] [--profile]

Builds scripts/null_device/nulllib.c (every C-ABI entry point returns success at once) into a
scratch copy of the package under /tmp and runs BASELINE config 1 / config 2 / the C4 step there,
so the wall time per iteration is pure host work.  Used to drive VERDICT r1 item 7 (<= 5 us/op)."""
import cProfile
import os
import pstats
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRATCH = "/tmp/mdb_null_device"


def prepare():
    pkg = os.path.join(SCRATCH, "minidiff_b200")
    shutil.rmtree(SCRATCH, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "minidiff_b200"), pkg,
                    ignore=shutil.ignore_patterns("csrc", "lib", "__pycache__"))
    os.makedirs(os.path.join(pkg, "lib"))
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", os.path.join(pkg, "lib", "libminidiff_b200.so"),
                    os.path.join(ROOT, "scripts", "null_device", "nulllib.c")], check=True)
    sys.path.insert(0, SCRATCH)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c1"
    profile = "--profile" in sys.argv
    sys.argv = sys.argv[:1]
    prepare()
    import numpy as np

    import minidiff_b200 as md
    from minidiff_b200.backend._lib import lib

    assert md.__file__.startswith(SCRATCH)
    if which == "c1":
        x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
        y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)

        def step():
            f = 2 * y * md.sin(x) - x ** 2
            f.backward(allow_higher_order=True)
            x.grad.backward()
    elif which == "c2":
        a = md.Tensor(np.zeros((8192, 1), np.float32), allow_grad=True)
        c = md.Tensor(np.zeros((1, 8192), np.float32), allow_grad=True)

        def step():
            md.sum(md.sin(a * c + a) ** 2).backward()
    else:
        from minidiff_b200 import workloads as W

        params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
        X, Y = md.Tensor(np.zeros((8192, 1024), np.float32)), md.Tensor(np.zeros((8192, 1024), np.float32))

        def step():
            W.mlp_train_step(X, Y, params)
    for _ in range(200):
        step()
    n = 2000
    l0 = lib.mdb_launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t0
    per = (lib.mdb_launch_count() - l0) / n
    print(f"{which}: {dt / n * 1e6:.1f} us per iteration, {per:.0f} launches -> {dt / n * 1e6 / per:.2f} us per launch (host only)")
    if profile:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(500):
            step()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
