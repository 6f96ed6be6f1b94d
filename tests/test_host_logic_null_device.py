"""Host-side logic of the engine WITHOUT a GPU: the package is copied to a scratch directory with a
NULL DEVICE in place of libminidiff_b200.so (scripts/null_device/nulllib.c: every C-ABI entry point
returns success at once, nothing is computed) and the BASELINE workloads are driven through it in a
subprocess.  What can be checked there is everything the host decides: how many launches a workload
issues (fusion, aliasing, in-place accumulation), result shapes / dtypes / strides against NumPy's,
error types, and that the null device exports every symbol of the ABI (so it cannot drift)."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

DRIVER = r'''
import json, sys
sys.path.insert(0, "scripts")
import host_profile
sys.argv = sys.argv[:1]
host_profile.prepare()
import numpy as np
import minidiff_b200 as md
from minidiff_b200.backend import _lib
from minidiff_b200 import workloads as W
assert md.__file__.startswith(host_profile.SCRATCH)
L = lambda: int(_lib.lib.mdb_launch_count())
out = {"exports": len(_lib.EXPORTS)}
x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)
l0 = L(); f = 2 * y * md.sin(x) - x ** 2; out["c1_forward"] = L() - l0
l0 = L(); f.backward(allow_higher_order=True); out["c1_backward_recorded"] = L() - l0
l0 = L(); x.grad.backward(); out["c1_second_order"] = L() - l0
out["c1_meta"] = [list(f.shape), str(f.dtype), list(x.grad.shape), str(x.grad.dtype)]
a = md.Tensor(np.zeros((64, 1), np.float32), allow_grad=True); c = md.Tensor(np.zeros((1, 48), np.float32), allow_grad=True)
l0 = L(); loss = md.sum(md.sin(a * c + a) ** 2); out["c2_forward"] = L() - l0
l0 = L(); loss.backward(); out["c2_backward_fused"] = L() - l0
out["c2_meta"] = [list(loss.shape), list(a.grad.shape), list(c.grad.shape)]
params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params((1024, 4096, 4096, 1024))]
X, Y = md.Tensor(np.zeros((512, 1024), np.float32)), md.Tensor(np.zeros((512, 1024), np.float32))
l0 = L(); W.mlp_train_step(X, Y, params); out["c4_step"] = L() - l0
l0 = L()
h = md.linear_relu(X, params[0], params[1]); h = md.linear_relu(h, params[2], params[3])
lf = md.mean((md.linear(h, params[4], params[5]) - Y) ** 2); lf.backward(); out["c4_fused_fwd_bwd"] = L() - l0
out["grad_shapes"] = [list(p.grad.shape) for p in params]
# metadata semantics against NumPy (no arithmetic involved)
B = md.backend
t = B.zeros((3, 1, 5), dtype=np.float32); u = B.zeros((4, 1), dtype=np.float64)
meta = {}
meta["bcast"] = [list((t + u).shape), str((t + u).dtype)]
meta["scalar_weak"] = [str((t * 2).dtype), str((t * 2.5).dtype), str((B.zeros((2,), dtype=np.int64) * 2.5).dtype)]
meta["cmp"] = str((t > 0).dtype)
meta["sum"] = [list(B.sum(t, axis=(0, 2)).shape), list(B.sum(t, axis=1, keepdims=True).shape), list(B.sum(t).shape)]
meta["T"] = [list(t.T.shape), list(t.T.strides)]
meta["reshape_view"] = t.reshape(3, 5)._st is t._st
meta["flip"] = list(B.flip(B.zeros((4, 6), dtype=np.float32), 1).strides)
meta["matmul"] = [list(B.matmul(B.zeros((7, 3), dtype=np.float32), B.zeros((3, 2), dtype=np.float32)).shape),
                  list(B.matmul(B.zeros((5, 7, 3)), B.zeros((3, 2))).shape), str(B.matmul(B.zeros((2, 2), dtype=np.int64), B.zeros((2, 2), dtype=np.int64)).dtype)]
errs = {}
for name, fn in (("bcast", lambda: t + B.zeros((2, 4), dtype=np.float32)), ("matmul", lambda: B.matmul(B.zeros((2, 3)), B.zeros((4, 2)))),
                 ("readonly", lambda: B.broadcast_to(t, (2, 3, 1, 5)).__iadd__(1)), ("axis", lambda: B.sum(t, axis=3)),
                 ("index", lambda: t[5]), ("cast", lambda: B.zeros((2,), dtype=np.int64).__iadd__(1.5))):
    try:
        fn(); errs[name] = "no error"
    except Exception as exc:
        errs[name] = type(exc).__name__
out["meta"], out["errors"] = meta, errs
print("RESULT " + json.dumps(out))
'''


def run_driver():
    r = subprocess.run([sys.executable, "-c", DRIVER], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_null_device_exports_the_whole_abi():
    import re

    header = open(os.path.join(ROOT, "include", "minidiff_b200.h")).read()
    declared = set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", header))
    null_src = open(os.path.join(ROOT, "scripts", "null_device", "nulllib.c")).read()
    defined = set(re.findall(r"\b(mdb_[a-z0-9_]+)\b", null_src))
    assert declared <= defined, sorted(declared - defined)


def test_host_logic_launch_counts_shapes_and_errors():
    got = run_driver()
    # launches the host issues per workload (a change here is a change of the fusion / aliasing logic)
    assert got["c1_forward"] == 5
    assert got["c1_forward"] + got["c1_backward_recorded"] + got["c1_second_order"] == 29
    # reference chain: 8 forward + 14 backward calls; fused backward here: seed + 5 fused gradient launches
    # (at 8192 x 8192 the column sum splits its rows and adds one fold launch: 12 per iteration)
    assert got["c2_forward"] == 5 and got["c2_backward_fused"] == 6
    assert got["c4_step"] == 38       # 41 at batch 65536: the three bias-gradient column sums add a fold launch each
    assert got["c4_fused_fwd_bwd"] <= 20
    assert got["c1_meta"] == [[2, 4], "float32", [2, 4], "float32"]
    assert got["c2_meta"] == [[], [64, 1], [1, 48]]
    assert got["grad_shapes"] == [[1024, 4096], [4096], [4096, 4096], [4096], [4096, 1024], [1024]]
    m = got["meta"]
    a, b = np.zeros((3, 1, 5), np.float32), np.zeros((4, 1))
    assert m["bcast"] == [list((a + b).shape), str((a + b).dtype)]
    assert m["scalar_weak"] == [str((a * 2).dtype), str((a * 2.5).dtype), str((np.zeros(2, np.int64) * 2.5).dtype)]
    assert m["cmp"] == "bool"
    assert m["sum"] == [list(a.sum(axis=(0, 2)).shape), list(a.sum(axis=1, keepdims=True).shape), []]
    assert m["T"][0] == list(a.T.shape) and m["T"][1][0] == a.T.strides[0]
    assert m["reshape_view"] is True
    assert m["flip"] == list(np.zeros((4, 6), np.float32)[:, ::-1].strides)
    assert m["matmul"] == [[7, 2], [5, 7, 2], "int64"]
    assert got["errors"] == {"bcast": "ValueError", "matmul": "ValueError", "readonly": "ValueError",
                             "axis": "AxisError", "index": "IndexError", "cast": "TypeError"}
