"""Data-parallel gradient exchange for the MLP training step (BASELINE config 4).

One process per GPU.  Parameters are replicated, the batch is sharded by rows; after the local
backward each parameter gradient is averaged over ranks with NCCL (over NVLink 5 / NVSwitch).
The all-reduce of a parameter is issued on the library's comm stream the moment the backward
sweep has finished that parameter's gradient (Tensor grad-ready hook), so the exchange of the
late layers overlaps the remaining backward GEMMs; `finish()` makes the compute stream wait for
the comm stream before the optimiser touches the gradients.

The reference has no distributed code; this module is new.  Rendezvous (passing the
NCCL unique id) uses torch.distributed's gloo group: plumbing only, no tensor goes through torch.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

from minidiff_b200.backend import _lib
from minidiff_b200.backend._lib import check, lib


def find_nccl() -> str:
    try:
        import nvidia.nccl as n

        for base in list(getattr(n, "__path__", [])):
            hits = glob.glob(os.path.join(base, "lib", "libnccl.so*"))
            if hits:
                return hits[0]
    except Exception:
        pass
    return "libnccl.so.2"


class DataParallel:
    def __init__(self, params, rank=None, world=None, overlap=True):
        import torch.distributed as dist

        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
        self.params = list(params)
        self.overlap = overlap
        self._pending = False
        self._seq = {}
        self._held = []
        self._flushed = False
        if self.world == 1:
            return
        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
        _lib.ensure_device()
        path = find_nccl().encode()
        uid = (C.c_char * 128)()
        if self.rank == 0:
            check(lib.mdb_comm_unique_id(uid, path))
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0)
        # NCCL prints its version banner on stdout; keep stdout clean for callers that parse it
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            check(lib.mdb_comm_init(self.rank, self.world, box[0], path))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        if overlap:
            for p in self.params:
                p._grad_hook = self._on_grad_ready

    SMALL = 1 << 16          # gradients below this many elements wait for the next large one

    # called by the backward sweep as soon as p.grad is final
    def _on_grad_ready(self, p):
        """Bucketing without a flat arena: a small gradient (a bias: a few KB) is held back and rides
        in the NCCL launch of the next large one (its layer's weight gradient, which the backward sweep
        finishes right after) -- ncclGroupStart/End fuses the calls into one kernel, so the C4 step
        issues 3 exchanges instead of 6."""
        self._held.append(p)
        if p.grad._data.size >= self.SMALL:
            self._launch_held()

    def _private_f32(self, p):
        g = p.grad._data
        # all-reduce IN PLACE only when this sweep provably owns the buffer (a fresh GEMM / reduction
        # output).  The engine aliases gradients like the reference does -- `a + b` hands the upstream
        # gradient to both inputs, `reshape` returns it -- so p.grad can be the buffer of an
        # intermediate's gradient that the compute stream is still reading for the rest of backward:
        # averaging it in place on the comm stream would be a data race.  Give it private storage first.
        private = getattr(p, "_grad_private", None) is p.grad
        if not private or g.dtype != np.float32 or not g.is_c_contiguous() or not g.writeable:
            from minidiff_b200.backend import functions as F
            import minidiff_b200 as md

            p.grad = md.Tensor(F.astype(g, np.float32))
            g = p.grad._data
        return g

    def _launch_held(self):
        if not self._held:
            return
        bufs = [self._private_f32(p) for p in self._held]
        n = len(bufs)
        if n == 1:
            check(lib.mdb_comm_allreduce_f32(bufs[0].ptr, bufs[0].size, 1))
        else:
            ptrs = (C.c_void_p * n)(*[b.ptr for b in bufs])
            counts = (C.c_size_t * n)(*[b.size for b in bufs])
            check(lib.mdb_comm_allreduce_multi_f32(ptrs, counts, n, 1))
        seq = int(lib.mdb_comm_last_seq())
        for p in self._held:
            self._seq[id(p)] = seq
        self._held = []
        self._pending = True

    def _allreduce(self, p):
        self._held.append(p)
        self._launch_held()

    def flush(self):
        """Issue the exchanges that the backward hooks did not (overlap=False); no waiting."""
        if self.world > 1 and not self.overlap and not self._flushed:
            self._held = list(self.params)
            self._flushed = True
        if self.world > 1:
            self._launch_held()          # small gradients still waiting for a large one

    def wait(self, p):
        """Order the compute stream after the exchange of THIS parameter's gradient only, so the
        optimiser can update early layers' parameters while later exchanges are still in flight."""
        if self.world == 1:
            return
        seq = self._seq.pop(id(p), 0)
        if seq:
            check(lib.mdb_comm_wait_seq(seq))

    def update_order(self):
        """Parameters in the order their gradients finish in the backward sweep (last layer first)."""
        return list(reversed(self.params))

    def finish(self):
        """All gradients averaged and visible to the compute stream after this returns (async)."""
        if self.world == 1:
            return
        self.flush()
        if self._pending:
            check(lib.mdb_comm_wait())
            self._pending = False
        self._seq.clear()
        self._flushed = False

    def close(self):
        if self.world > 1:
            for p in self.params:
                p._grad_hook = None
            check(lib.mdb_comm_destroy())
