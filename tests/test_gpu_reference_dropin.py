"""The drop-in claim on real hardware: the UNMODIFIED reference package (installed once, offline, with
`pip install --target baseline/_ref /root/reference`; git-ignored but shipped to the GPU box) selects
`minidiff_b200.plugin` through its own `--backend` loader and runs its own Tensor / create_op_func /
OpNode code -- forward, first- and second-order backward -- on device storage.  Results are compared
with the committed golden vectors (generated from the reference's NumPy backend) and with the oracle.
Skipped when baseline/_ref is absent (a checkout without the one-off install)."""
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")

SCRIPT = r"""
import sys, os
sys.argv = [sys.argv[0], "--backend", "minidiff_b200.plugin"]
import numpy as np
import minidiff as md                      # the reference package itself
import minidiff_b200.plugin as plugin
plugin.assert_live(md)
import minidiff.backend as live
from minidiff_b200.backend.device_array import DeviceArray
assert live.tensor_class is DeviceArray
assert md.Tensor.__module__ == "minidiff.tensor" and "baseline/_ref" in md.__file__.replace(os.sep, "/")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_minidiff as orc

def host(t):
    return t._data.numpy() if hasattr(t._data, "numpy") else np.asarray(t._data)

# ---- C1: README example, first + second order (golden from the reference's NumPy backend)
g = np.load(os.path.join(GOLDEN, "c1.npz"))
x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)
f = 2 * y * md.sin(x) - x**2
f.backward(allow_higher_order=True)
np.testing.assert_allclose(host(f), g["f"], rtol=1e-6, atol=1e-6)
np.testing.assert_allclose(host(x.grad), g["dx"], rtol=1e-6, atol=1e-6)
np.testing.assert_allclose(host(y.grad), g["dy"], rtol=1e-6, atol=1e-6)
x.grad.backward()
np.testing.assert_allclose(host(x.grad), g["dxx"], rtol=1e-6, atol=1e-6)
np.testing.assert_allclose(host(y.grad), g["dxy"], rtol=1e-6, atol=1e-6)

# ---- C2: broadcast chain with un-broadcast gradient sums (reference unbroadcast path on the device)
a_np, c_np = orc.config2_inputs(257, 131)
w2 = orc.config2(a_np, c_np)
a, c = md.Tensor(a_np, allow_grad=True), md.Tensor(c_np, allow_grad=True)
loss = md.sum(md.sin(a * c + a) ** 2)
loss.backward()
np.testing.assert_allclose(host(loss), w2["loss"], rtol=1e-4)
np.testing.assert_allclose(host(a.grad), w2["da"], rtol=1e-4, atol=1e-3)
np.testing.assert_allclose(host(c.grad), w2["dc"], rtol=1e-4, atol=1e-3)

# ---- C4: MLP training step (matmul -> tcgen05 GEMMs, where-ReLU, mean-MSE, in-place SGD)
dims, B = (128, 256, 256, 128), 512
X, Y = orc.mlp_data(B, dims[0], dims[-1])
ps_np = orc.mlp_params(dims)
w4 = orc.config4_step(X, Y, ps_np)
ps = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
h = md.Tensor(X)
for l in range(3):
    h = h @ ps[2 * l] + ps[2 * l + 1]
    if l < 2:
        h = md.where(h > 0, h, 0)
mse = md.mean((h - md.Tensor(Y)) ** 2)
mse.backward()
np.testing.assert_allclose(host(mse), w4["loss"], rtol=1e-4)
for p, gr in zip(ps, w4["grads"]):
    np.testing.assert_allclose(host(p.grad), gr, rtol=1e-4, atol=1e-6)
with md.no_grad():
    for p in ps:
        p -= 0.01 * p.grad
for p, want in zip(ps, w4["params"]):
    np.testing.assert_allclose(host(p), want, rtol=1e-4, atol=1e-6)
from minidiff_b200.backend._lib import lib
print("DROPIN_OK launches", lib.mdb_launch_count())
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "minidiff")),
                    reason="baseline/_ref (offline install of the reference) not present")
def test_unmodified_reference_runs_on_the_device_through_the_plugin():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, os.path.join(ROOT, "oracle", "_stubs"), ROOT])
    code = f"ROOT = {ROOT!r}\nGOLDEN = {GOLDEN!r}\n" + SCRIPT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
