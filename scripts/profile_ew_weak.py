"""The elementwise / reduction cases that sit lowest against the HBM roofline, one launch each
(for ncu --set full)."""
import ctypes as C, sys
sys.path.insert(0, ".")
sys.argv = sys.argv[:1]
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend._lib import lib, check
from minidiff_b200.backend import functions as F
N = M = 8192
rng = np.random.default_rng(0)
t = B.asarray(rng.standard_normal((N, M), dtype=np.float32))
a = B.asarray(rng.standard_normal((N, 1), dtype=np.float32))
c = B.asarray(rng.standard_normal((1, M), dtype=np.float32))
out_a = B.zeros((N, 1), dtype=np.float32)
out_c = B.zeros((1, M), dtype=np.float32)


def ered(op, out, *ins, acc=0):
    n = len(ins)
    descs = (F.MdbArray * n)()
    for i, o in enumerate(ins):
        descs[i] = o.d
    check(lib.mdb_elementwise_reduce(F.OP[op], C.byref(out.d), n, descs, acc))


for rep in range(2):
    B.multiply(a, c)            # outer
    B.sum(t, axis=1)
    ered("MUL", out_a, t, c)
    ered("MUL", out_c, t, a)
    B.synchronize()
print("done")
