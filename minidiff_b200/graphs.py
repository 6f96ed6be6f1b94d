"""CUDA-graph capture of a repeating step.

The reference's `caching.reuse_graph` (caching.py:15-65) tells the engine "the same graph is built
again and again" and saves the host-side toposort.  On the device the same promise is worth much
more: the whole step -- forward ops, the backward sweep with its fused accumulate kernels, the
in-place SGD update -- is captured once into a CUDA graph and replayed with a single launch, which
removes the per-op Python / ctypes / launch cost (what dominates BASELINE config 1 and small-batch
steps).

    step = md.capture_graph(lambda: train_step(X, Y, params))   # 2 eager warm-up calls, then ONE
                                                                 # recorded (not executed) call
    for _ in range(n):
        X[...] = next_batch          # refresh inputs IN PLACE (device copy), addresses stay fixed
                                     # (under md.no_grad() if X itself tracks gradients)
        loss = step.replay()         # the tensors the function returned, updated in place

Rules inside the captured function (checked by the C layer where possible): no `.item()`,
`.as_numpy()`, `print` of tensors or other read-backs, no creation of tensors from host data, no
data-dependent Python control flow, profiler off.  Data-parallel steps (NCCL on the comm stream) are
not captured.
"""
from __future__ import annotations

import ctypes as C

from .backend._lib import check, lib


class CapturedGraph:
    def __init__(self, fn, *args, warmup: int = 2, **kwargs):
        # warm-up: first-launch setup (kernel attributes, tensor-map driver entry point) and the
        # allocator's steady state must exist before the stream goes into capture mode
        for _ in range(max(0, warmup)):
            fn(*args, **kwargs)
        check(lib.mdb_sync())
        self._handle = C.c_void_p()
        check(lib.mdb_graph_begin())
        try:
            self.result = fn(*args, **kwargs)
        except BaseException:
            junk = C.c_void_p()
            if lib.mdb_graph_end(C.byref(junk)) == 0:
                lib.mdb_graph_destroy(junk)
            raise
        check(lib.mdb_graph_end(C.byref(self._handle)))
        n, b = C.c_uint64(), C.c_size_t()
        check(lib.mdb_graph_info(self._handle, C.byref(n), C.byref(b)))
        self.kernel_launches, self.pinned_bytes = int(n.value), int(b.value)
        # the capture only RECORDED the step: `result` (and e.g. the parameters' .grad tensors) hold
        # meaningful values after the first replay()

    def replay(self):
        """Run the captured step once more (asynchronous); returns what the function returned."""
        if not self._handle:
            raise RuntimeError("this graph was closed")
        check(lib.mdb_graph_launch(self._handle))
        return self.result

    __call__ = replay

    def close(self):
        if self._handle:
            check(lib.mdb_graph_destroy(self._handle))
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def capture_graph(fn, *args, warmup: int = 2, **kwargs) -> CapturedGraph:
    return CapturedGraph(fn, *args, warmup=warmup, **kwargs)
