"""Interleaved A/B of the C4 training step: this repo's engine vs the UNMODIFIED reference engine on the
same device backend (same process, alternating blocks of steps, so both see the same clocks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in (os.path.join(ROOT, "oracle", "_stubs"), os.path.join(ROOT, "baseline", "_ref")):
    sys.path.insert(0, p)
sys.argv = [sys.argv[0], "--backend", "minidiff_b200.plugin"]
import minidiff as ref
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200 as md
import minidiff_b200.plugin as plugin
from minidiff_b200 import workloads as W
from bench import Dev, DIMS, LR

plugin.assert_live(ref)
dev = Dev()
B = 65536
X_np, Y_np = W.mlp_data(B, DIMS[0], DIMS[-1], seed=1000)


def make(m):
    params = [m.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]
    X, Y = m.Tensor(X_np), m.Tensor(Y_np)

    def step():
        h = X
        for l in range(3):
            h = h @ params[2 * l] + params[2 * l + 1]
            if l < 2:
                h = m.where(h > 0, h, 0)
        loss = m.mean((h - Y) ** 2)
        loss.backward()
        with m.no_grad():
            for p in params:
                p -= LR * p.grad
        return loss
    return step


steps = {"b200_engine": make(md), "reference_engine": make(ref)}
for s in steps.values():
    for _ in range(4):
        s()
dev.sync()
res = {k: [] for k in steps}
e0, e1 = dev.event(), dev.event()
for rnd in range(6):
    for name, s in steps.items():
        l0 = dev.launches()
        dev.record(e0)
        for _ in range(5):
            s()
        dev.record(e1)
        dev.sync()
        res[name].append(dev.elapsed_ms(e0, e1) / 5)
        launches = (dev.launches() - l0) / 5
    print("round", rnd, {k: round(v[-1], 3) for k, v in res.items()}, flush=True)
for k, v in res.items():
    print(f"{k:18s} median {np.median(v):.3f} ms/step  min {min(v):.3f}")
