"""Host (Python + ctypes) time per C4 training step vs device time, at the per-rank batch of an
8-GPU data-parallel run (8192 rows) and at a tiny batch where the step is purely launch-bound.
    python scripts/host_overhead.py"""
import sys, time
sys.path.insert(0, ".")
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200 as md
from minidiff_b200 import workloads as W
from bench import Dev

dev = Dev()
for batch in (8192, 256):
    X_np, Y_np = W.mlp_data(batch, 1024, 1024)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    for _ in range(5):
        W.mlp_train_step(X, Y, params)
    dev.sync()
    e0, e1 = dev.event(), dev.event()
    n = 30
    l0 = dev.launches()
    t0 = time.perf_counter()
    dev.record(e0)
    for _ in range(n):
        W.mlp_train_step(X, Y, params)
    t_issue = time.perf_counter() - t0
    dev.record(e1)
    dev.sync()
    t_wall = time.perf_counter() - t0
    print(f"batch {batch}: host issue {t_issue/n*1e3:.3f} ms/step, wall {t_wall/n*1e3:.3f} ms/step, "
          f"device {dev.elapsed_ms(e0, e1)/n:.3f} ms/step, launches/step {(dev.launches()-l0)/n:.0f}", flush=True)
# the same step as ONE CUDA-graph replay (no per-op host work, back-to-back kernels)
for batch in (8192,):
    X_np, Y_np = W.mlp_data(batch, 1024, 1024)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    g = md.capture_graph(lambda: W.mlp_train_step(X, Y, params))
    for _ in range(3):
        g.replay()
    dev.sync()
    e0, e1 = dev.event(), dev.event()
    dev.record(e0)
    for _ in range(30):
        g.replay()
    dev.record(e1)
    dev.sync()
    print(f"batch {batch}: CUDA-graph replay device {dev.elapsed_ms(e0, e1)/30:.3f} ms/step ({g.kernel_launches} launches)", flush=True)
    g.close()
    # per-class device time of the eager step
    for _ in range(3):
        W.mlp_train_step(X, Y, params)
    dev.prof(True)
    for _ in range(10):
        W.mlp_train_step(X, Y, params)
    dev.sync()
    for cls, name in ((0, "elementwise"), (1, "reduce"), (2, "gemm")):
        ms, n, w = dev.prof_read(cls)
        print(f"   {name:12s} {ms/10:.3f} ms/step in {n/10:.0f} calls, {w/ms/1e6 if cls < 2 else w/ms/1e9:.0f} {'GB/s' if cls < 2 else 'TFLOP/s'}")
    dev.prof(False)
import cProfile, pstats
X_np, Y_np = W.mlp_data(256, 1024, 1024)
params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
X, Y = md.Tensor(X_np), md.Tensor(Y_np)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    W.mlp_train_step(X, Y, params)
pr.disable()
dev.sync()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
