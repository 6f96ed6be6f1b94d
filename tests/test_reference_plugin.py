"""Drop-in check against the REAL reference package (build container only: /root/reference does
not travel to the GPU box).  The reference's own loader must pick up minidiff_b200.plugin through
`--backend`, export the 114 names into `minidiff.backend`, and bind its unchanged ops/definitions
to the device functions."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("MINIDIFF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "minidiff")),
                                reason="reference checkout not present")

SCRIPT = r"""
import sys
sys.argv = [sys.argv[0], "--backend", "minidiff_b200.plugin"]
import minidiff as md
import minidiff.backend as live
import minidiff_b200.plugin as plugin
from minidiff_b200.backend.device_array import DeviceArray, c_strides
import numpy as np
plugin.assert_live(md)
assert live.tensor_class is DeviceArray
names = [n for n in vars(plugin.b200_backend) if not n.startswith("_")]
assert len(names) == 114, len(names)
assert live.multiply is plugin.b200_backend.multiply
# the reference's unchanged Tensor / op wrapper / OpNode run on device storage (metadata-only here)
raw = DeviceArray(None, 1 << 20, (2, 3), c_strides((2, 3)), np.dtype(np.float32))
t = md.Tensor(raw, allow_grad=True)
v = md.transpose(t)
assert v.shape == (3, 2) and v.op_node.op_name == "transpose" and t.graph_refs == 1
u = md.reshape(t, (6,))
assert u.shape == (6,)
print("PLUGIN_OK", len(names))
"""


def test_reference_loader_selects_the_b200_plugin():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, os.path.join(ROOT, "oracle", "_stubs"), ROOT])
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "PLUGIN_OK 114" in r.stdout, r.stdout + r.stderr
