"""Worker for the data-parallel tests (launched with torch.distributed.run, one rank per process).

mode "gloo-oracle": CPU only.  Each rank runs the ORACLE's backward on its row shard, gradients are
    averaged with a gloo all-reduce -- the exact exchange parallel.DataParallel performs with NCCL --
    and rank 0 checks them against the oracle's full-batch gradients.
mode "nccl": one GPU per rank through minidiff_b200.parallel.DataParallel; rank 0 saves the updated
    parameters for comparison with a single-GPU step on the full batch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
mode, out = sys.argv[1], sys.argv[2]
sys.argv = sys.argv[:1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
DIMS, BATCH = (32, 64, 64, 16), 256

import np_minidiff as orc  # noqa: E402

X, Y = orc.mlp_data(BATCH, DIMS[0], DIMS[-1])
ps_np = orc.mlp_params(DIMS)
lo, hi = rank * BATCH // world, (rank + 1) * BATCH // world

if mode == "gloo-oracle":
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo")
    local = orc.config4_step(X[lo:hi], Y[lo:hi], ps_np)["grads"]
    flat = torch.from_numpy(np.concatenate([g.ravel() for g in local]).astype(np.float32))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    if rank == 0:
        full = orc.config4_step(X, Y, ps_np)["grads"]
        want = np.concatenate([g.ravel() for g in full])
        np.testing.assert_allclose(flat.numpy(), want, rtol=1e-4, atol=1e-6)
        open(out, "w").write("OK")
    dist.barrier()
    dist.destroy_process_group()
else:
    import minidiff_b200 as md
    from minidiff_b200 import workloads as W
    from minidiff_b200.parallel import DataParallel

    params = [md.Tensor(p.copy(), allow_grad=True) for p in ps_np]
    dp = DataParallel(params, rank, world)
    Xs, Ys = md.Tensor(X[lo:hi]), md.Tensor(Y[lo:hi])
    for _ in range(2):
        loss = W.mlp_train_step(Xs, Ys, params, 0.01, dp)
    # a parameter whose gradient ALIASES an intermediate's gradient (same-shape additive term: the
    # engine hands the upstream gradient buffer to both inputs of `+`, like the reference): the
    # exchange must not average that shared buffer in place while backward still reads it
    rng = np.random.default_rng(77)
    Wa_np = rng.standard_normal((DIMS[0], 1024)).astype(np.float32)
    Pa_np = rng.standard_normal((BATCH // world, 1024)).astype(np.float32)
    Wa, Pa = md.Tensor(Wa_np.copy(), allow_grad=True), md.Tensor(Pa_np.copy(), allow_grad=True)
    dp2 = DataParallel.__new__(DataParallel)          # second parameter set on the same communicator
    dp2.__dict__.update(dp.__dict__)
    dp2.params, dp2._seq, dp2._pending, dp2._flushed, dp2._held = [Wa, Pa], {}, False, False, []
    for q in dp2.params:
        q._grad_hook = dp2._on_grad_ready
    alias = []
    for _ in range(3):
        h = Xs @ Wa + Pa
        loss_a = md.mean(h ** 2)
        loss_a.backward()
        dp2.finish()
        alias = [Wa.grad.as_numpy(), Pa.grad.as_numpy()]
    if rank == 0:
        np.savez(out, *[p.as_numpy() for p in params], loss=loss.as_numpy(), alias_dW=alias[0], alias_dP=alias[1])
    import torch.distributed as dist

    dist.barrier()
    dp.close()
    dist.destroy_process_group()
