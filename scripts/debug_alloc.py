import sys, ctypes as C
sys.path.insert(0, ".")
import minidiff_b200 as md
from minidiff_b200 import workloads as W
from minidiff_b200.backend._lib import lib
B = 65536
X_np, Y_np = W.mlp_data(B, 1024, 1024)
params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
X, Y = md.Tensor(X_np), md.Tensor(Y_np)
def allocs():
    v = [C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_uint64()]
    lib.mdb_mem_stats(*[C.byref(x) for x in v]); return v[3].value, v[0].value/1e9, v[1].value/1e9
for i in range(8):
    print("---- step", i, allocs(), file=sys.stderr, flush=True)
    loss = W.mlp_train_step(X, Y, params)
    md.backend.synchronize()
print("end", allocs(), file=sys.stderr)
