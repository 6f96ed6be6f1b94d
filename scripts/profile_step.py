"""One short pass over the BASELINE workloads for ncu (launch list / --set full captures).
    python scripts/profile_step.py [mlp|c2|c3|all]
"""
import sys
sys.path.insert(0, ".")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
sys.argv = sys.argv[:1]
import minidiff_b200 as md
from minidiff_b200 import workloads as W

md.backend.assert_live()
if which in ("mlp", "all"):
    B = 65536
    X_np, Y_np = W.mlp_data(B, 1024, 1024)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params()]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    for _ in range(2):
        loss = W.mlp_train_step(X, Y, params)
    print("mlp loss", loss.item())
    del X, Y, params, loss
if which in ("c2", "all"):
    a_np, c_np = W.c2_inputs()
    a, c = md.Tensor(a_np, allow_grad=True), md.Tensor(c_np, allow_grad=True)
    for _ in range(2):
        loss = W.c2_step(a, c)
    print("c2 loss", loss.item())
    del a, c, loss
if which in ("c3", "all"):
    A_np, B_np = W.c3_inputs(8192)
    A, Bm = md.Tensor(A_np, allow_grad=True), md.Tensor(B_np, allow_grad=True)
    for _ in range(2):
        C = W.c3_step(A, Bm)
    md.backend.synchronize()
    print("c3 done", C.shape)
