#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for minidiff_b200.

Headline line (one JSON object on stdout, rank 0):
  metric  = MLP train samples/s (BASELINE config 4: 3-layer MLP 1024-4096-4096-1024, where-ReLU,
            mean-MSE, SGD; global batch 65536 fp32 synthetic; data-parallel over N GPUs with NCCL
            all-reduce of the parameter gradients).  A "step" is one full training step.
  value   = whole-job samples/s with inputs resident in HBM;   e2e = same through the public API
            with pinned HOST inputs (H2D of X,Y and D2H of the loss inside the timed region).
  roofline       = the dominant kernel class of the step (GEMM), timed live with CUDA events on the
                   library's launch stream during the timed region.  ONE tensor denominator everywhere:
                   MEASURED_PEAKS.json bf16_tflops_sustained / 2 (all GEMM legs are timed inside
                   multi-launch loops under the power cap); frac_of_burst uses bf16_tflops / 2.  The
                   cuBLAS TF32 rate measured in this run is reported as a side note only.
                   pipe_frac = 3 * frac: a 3xTF32 GEMM issues three tensor-core MACs per fp32 product.
  cpu_baseline   = the UNMODIFIED reference (baseline/_ref, its NumPy backend) on this box's host
                   cores, same workload, bounded number of steps; also inside fwd_bwd.c2/c3/c5.
  fwd_bwd (N=1)  = the single-GPU graph benchmarks of BASELINE.json: config 1 (README example, eager
                   vs one CUDA-graph replay, microseconds), config 2 (broadcast chain, GB/s vs HBM
                   roofline), config 3 (matmul fwd+bwd) and config 5 (Hessian-vector product), TFLOP/s
                   vs the tensor roofline; c4_reference_engine_dropin = the same C4 step run by the
                   UNMODIFIED reference engine with only --backend switched; c4_fused = the step
                   written with md.linear_relu / md.linear (stateful fused ops).
  dp_parity (N>1) = before timing, the NCCL-averaged gradients of one data-parallel backward are
                   compared on rank 0 with a single-GPU backward over the same GLOBAL batch.

`--impl reference` runs the unmodified reference (baseline/_ref, NumPy backend, all host threads)
on the same workload and full batch; `--workload c2|c3|c5` selects the other configs (used by the
cpu_baseline legs).  Falls back to the oracle port only when baseline/_ref is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ncu --set full capture of one C4 step at N=1 (profiles/r02_ncu_mlp_step_gemm.md): DRAM bytes per GEMM
# launch, and the algorithmic figure beside it (each operand read once + C written once, 8 GEMMs/step)
GEMM_TRAFFIC_BYTES_PER_LAUNCH = 3.656e9
GEMM_ALGORITHMIC_BYTES_PER_LAUNCH = (
    # fwd1, fwd2, fwd3 (X@W), dW3, dh2, dW2, dh1, dW1 at B=65536, D=(1024,4096,4096,1024)
    sum(4.0 * (m * k + k * n + m * n) for m, k, n in [
        (65536, 1024, 4096), (65536, 4096, 4096), (65536, 4096, 1024),
        (4096, 65536, 1024), (65536, 1024, 4096), (4096, 65536, 4096), (65536, 4096, 4096),
        (1024, 65536, 4096)]) / 8.0)
# C2: DRAM bytes of the 13 launches of one iteration (ncu, profiles/r01_c2_launches_v3.csv)
C2_TRAFFIC_BYTES_PER_ITER = 3.990e9
C2_TRAFFIC_SRC = ("dram__bytes_read.sum + dram__bytes_write.sum over the 13 launches of one iteration, "
                  "profiles/r01_c2_launches_v3.csv (below the algorithmic 4.295e9: part of each output is still "
                  "in the 126 MB L2 when the next op reads it)")
GEMM_TRAFFIC_SRC = ("dram__bytes_read.sum + dram__bytes_write.sum averaged over the 8 GEMM launches of one C4 step "
                    "(ncu --set full, profiles/r02_ncu_mlp_step_gemm.md; captures of the same code on different "
                    "boxes gave 2.85-3.68 GB: what the L2 keeps between waves varies)")
C3_TRAFFIC_BYTES_PER_LAUNCH = 3.055e9
C3_TRAFFIC_SRC = ("dram__bytes_read.sum + dram__bytes_write.sum per 8192^3 launch of the shipped pair kernel "
                  "(ncu --set full, profiles/r02_ncu_c3_gemm.md)")
GLOBAL_BATCH = int(os.environ.get("MDB_BENCH_GLOBAL_BATCH", 65536))   # override for experiments only (scripts/dp_contention.sh)
DIMS = (1024, 4096, 4096, 1024)
LR = 0.01
METRIC = "mlp_train_samples_per_s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def tensor_roofline(tflops, peaks, **extra):
    """ONE denominator for every GEMM leg: MEASURED_PEAKS.json bf16_tflops_sustained / 2 (the legs
    are timed inside multi-launch loops, i.e. under the power cap); the burst figure and the cuBLAS
    TF32 rate measured in this run are side notes.  pipe_frac = 3 * frac (3xTF32 issues three
    tensor-core MACs per fp32 product)."""
    peak, burst = peaks["bf16_sustained"] / 2.0, peaks["bf16_burst"] / 2.0
    r = {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
         "frac": tflops / peak, "pipe_frac": 3.0 * tflops / peak, "frac_of_burst": tflops / burst,
         "peak_src": f"{peaks['src']} bf16_tflops_sustained / 2 (MEASURED_PEAKS.json: dense bf16 cuBLAS, "
                     "sustained under the power cap; TF32 runs at half the bf16 rate)",
         "note": "achieved = algorithmic 2MNK flops / CUDA-event time of the GEMM launches inside the timed "
                 "region; a 3xTF32 GEMM issues 3 tensor-core MACs per fp32 product, so the tensor pipe does "
                 "pipe_frac = 3*frac of the yardstick's work"}
    r.update(extra)
    return r


def measure_tf32_peak(seconds=2.0, n=8192):
    """TF32 dense tensor peak of THIS box, measured the way MEASURED_PEAKS.json measures bf16: a
    cuBLAS GEMM (torch.matmul with allow_tf32) at 8192^3, best of 10 (burst) and back to back for
    `seconds` (sustained, i.e. under the power cap).  SURVEY 8(d): the TF32 peak is not in
    MEASURED_PEAKS.json and has to be measured on the box.  Returns None if torch has no CUDA."""
    try:
        import torch

        if not torch.cuda.is_available():
            return None
        torch.backends.cuda.matmul.allow_tf32 = True
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        a = torch.randn(n, n, device=dev, dtype=torch.float32)
        b = torch.randn(n, n, device=dev, dtype=torch.float32)
        c = torch.empty(n, n, device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        flops = 2.0 * n ** 3
        best = None
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        t0 = time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                torch.matmul(a, b, out=c)
            reps += 20
            torch.cuda.synchronize(dev)
        e1.record(); e1.synchronize()
        sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b, c
        torch.cuda.empty_cache()
        return {"burst": flops / (best * 1e-3) / 1e12, "sustained": sustained,
                "how": f"torch.matmul fp32 with allow_tf32 (cuBLAS TF32) {n}^3: best of 10 and back to back for {seconds:.0f} s"}
    except Exception as exc:      # the bench must not die because the yardstick could not be taken
        return {"error": repr(exc)}


# ------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        for _ in range(20):                      # a very short timed region: wait for a sample taken after it began
            if any(x[0] >= (self.t0 or 0) for x in self.lines):
                break
            time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        lines = self.lines
        if self.t0 is not None and self.t1 is not None:
            # samples taken while the timed region ran (the sampler itself starts before warm-up so
            # that short regions still get samples); fall back to the nearest ones if none landed
            inside = [x for x in lines if self.t0 <= x[0] <= self.t1 + 0.06]
            lines = inside or sorted(lines, key=lambda x: abs(x[0] - self.t1))[:2]
        for _, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# helpers on top of the C ABI
# ------------------------------------------------------------------------------------------------
class Dev:
    def __init__(self):
        import ctypes as C

        import minidiff_b200 as md
        from minidiff_b200.backend import _lib

        self.C, self.md, self.lib, self.check = C, md, _lib.lib, _lib.check
        md.backend.assert_live()

    def event(self):
        e = self.C.c_void_p()
        self.check(self.lib.mdb_event_create(self.C.byref(e)))
        return e

    def record(self, e):
        self.check(self.lib.mdb_event_record(e))

    def elapsed_ms(self, a, b):
        ms = self.C.c_float()
        self.check(self.lib.mdb_event_elapsed_ms(a, b, self.C.byref(ms)))
        return ms.value

    def sync(self):
        self.check(self.lib.mdb_sync())

    def launches(self):
        return int(self.lib.mdb_launch_count())

    def mem(self):
        C = self.C
        v = [C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_uint64()]
        self.lib.mdb_mem_stats(*[C.byref(x) for x in v])
        return {"in_use_GB": v[0].value / 1e9, "cached_GB": v[1].value / 1e9, "peak_GB": v[2].value / 1e9,
                "device_allocs": v[3].value}

    def gemm_paths(self, reset=False):
        c = (self.C.c_uint64 * 8)()
        self.check(self.lib.mdb_gemm_stats(c, 1 if reset else 0))
        return {"simt": int(c[0]), "tc_single": int(c[1]), "tc_presplit": int(c[2]), "tc_pair": int(c[3]),
                "tc_pair_streamk": int(c[4])}

    def prof(self, on):
        self.check(self.lib.mdb_prof_enable(1 if on else 0))

    def prof_read(self, cls):
        C = self.C
        ms, n, w = C.c_double(), C.c_uint64(), C.c_double()
        self.check(self.lib.mdb_prof_read(cls, C.byref(ms), C.byref(n), C.byref(w)))
        return ms.value, n.value, w.value

    def pinned(self, arr):
        """copy a NumPy array into pinned host memory; returns (ndarray view, keepalive)"""
        C = self.C
        p = C.c_void_p()
        self.check(self.lib.mdb_host_alloc(arr.nbytes, C.byref(p)))
        buf = (C.c_char * arr.nbytes).from_address(p.value)
        view = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
        view[...] = arr
        return view, p

    def upload_into(self, dst_tensor, pinned_view):
        self.check(self.lib.mdb_h2d(dst_tensor._data.ptr, pinned_view.ctypes.data, pinned_view.nbytes))


def warm_until_stable(dev, fn, min_iters, max_iters=12):
    """Untimed warm-up: at least `min_iters` calls, then until one whole call makes no cudaMalloc
    (the caching allocator has every block size of this workload; a cudaMalloc inside a timed region
    is a device-wide synchronisation worth milliseconds)."""
    out = None
    for i in range(max_iters):
        before = dev.mem()["device_allocs"]
        out = fn()
        if i + 1 >= min_iters and dev.mem()["device_allocs"] == before:
            break
    dev.sync()
    return out


def dist_setup(world):
    if world == 1:
        return None
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo")
    return dist


def barrier(dist):
    if dist is not None:
        dist.barrier()


def all_ranks(dist, x):
    if dist is None:
        return [x]
    import torch

    out = [torch.zeros(1, dtype=torch.float64) for _ in range(dist.get_world_size())]
    dist.all_gather(out, torch.tensor([x], dtype=torch.float64))
    return [round(float(t.item()), 3) for t in out]


def max_over_ranks(dist, x):
    if dist is None:
        return x
    import torch

    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def bench_mlp(dev, dist, rank, world, steps, warmup, peaks, use_graph=False):
    graph_error = None
    md = dev.md
    from minidiff_b200 import workloads as W
    from minidiff_b200.parallel import DataParallel

    local = GLOBAL_BATCH // world
    # started first: nvidia-smi takes a few hundred ms to print its first sample on an 8-GPU box, and the
    # timed region of the 8-GPU run is only ~150 ms long
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", 0)))
    sampler.start()
    X_np, Y_np = W.mlp_data(local, DIMS[0], DIMS[-1], seed=1000 + 2 * rank)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    dp = DataParallel(params, rank, world) if world > 1 else None
    if os.environ.get("MDB_BENCH_GEMM_MAX_CLUSTERS"):          # experiment: leave SMs to the NCCL kernels
        dev.check(dev.lib.mdb_gemm_knob(7, int(os.environ["MDB_BENCH_GEMM_MAX_CLUSTERS"])))
    dp_parity = dp_parity_check(dev, dist, rank, world, dp, X, Y, params, local) if world > 1 else None

    def eager_step():
        return W.mlp_train_step(X, Y, params, LR, dp)

    loss = None
    for _ in range(warmup):
        loss = eager_step()      # same object lifetimes as the timed loop, so the allocator pool has converged
    dev.sync()
    # the repeating step as ONE CUDA-graph replay (md.capture_graph, the device-side counterpart of the
    # reference's caching.reuse_graph): forward, backward, the NCCL exchange on the forked comm stream
    # and the SGD update are captured once; every replay does the full work of a step
    graph, step = None, eager_step
    if use_graph:
        try:
            graph = md.capture_graph(eager_step, warmup=0)
            step = graph.replay
            for _ in range(2):
                loss = step()
            dev.sync()
        except Exception as exc:                      # capture refused (e.g. NCCL build without graph support)
            graph, step, graph_error = None, eager_step, repr(exc)
            for _ in range(2):
                loss = step()
            dev.sync()
    # ---- timed region: inputs resident in HBM
    e0, e1 = dev.event(), dev.event()
    if graph is None:
        dev.prof(True)           # per-class event timing brackets eager launches only
    barrier(dist)
    dev.sync()
    sampler.mark_begin()
    l0 = dev.launches()
    allocs0 = dev.mem()["device_allocs"]
    dev.gemm_paths(reset=True)
    dev.record(e0)
    for _ in range(steps):
        loss = step()
    dev.record(e1)
    dev.sync()
    sampler.mark_end()
    barrier(dist)
    ms = dev.elapsed_ms(e0, e1)
    launches = dev.launches() - l0
    mem = dev.mem()
    mem["device_allocs_in_timed_region"] = mem["device_allocs"] - allocs0
    clocks = sampler.stop()
    if graph is not None:
        # the per-class split (roofline of the GEMM class) comes from an eager pass of the same step right
        # after the timed region: CUDA events cannot bracket kernels inside a graph replay
        dev.gemm_paths(reset=True)
        dev.prof(True)
        for _ in range(min(steps, 5)):
            eager_step()
        dev.sync()
    gemm_ms, gemm_n, gemm_flops = dev.prof_read(2)
    ew_ms, ew_n, ew_bytes = dev.prof_read(0)
    red_ms, red_n, red_bytes = dev.prof_read(1)
    dev.prof(False)
    gemm_paths = dev.gemm_paths()
    prof_steps = steps if graph is None else min(steps, 5)
    rank_ms = all_ranks(dist, ms / steps)
    ms = max_over_ranks(dist, ms)
    loss_value = float(loss.item())

    # ---- e2e: host (pinned) inputs copied every step through the public input pipeline
    # (HostBatchFeeder: the upload of batch i+1 overlaps step i), loss read back every step
    feeder = W.HostBatchFeeder(X_np, Y_np)
    e2e_graphs = {}

    def e2e_step(Xd, Yd):
        """one training step on the device buffers the feeder just filled; in graph mode one captured
        graph per buffer pair of the double-buffered feeder (same work per replay as the eager step)"""
        if graph is None:
            return W.mlp_train_step(Xd, Yd, params, LR, dp)
        g = e2e_graphs.get(id(Xd))
        if g is None:
            g = e2e_graphs[id(Xd)] = md.capture_graph(lambda: W.mlp_train_step(Xd, Yd, params, LR, dp), warmup=0)
        return g.replay()

    last = None
    for _ in range(4):
        Xd, Yd = feeder.next()
        last = float(e2e_step(Xd, Yd).item())
    barrier(dist)
    dev.sync()
    t0 = time.perf_counter()
    f0, f1 = dev.event(), dev.event()
    dev.record(f0)
    for i in range(steps):
        Xd, Yd = feeder.next(prefetch_following=True)
        last = float(e2e_step(Xd, Yd).item())   # D2H read of the loss every step
    dev.record(f1)
    dev.sync()
    barrier(dist)
    e2e_ms = max_over_ranks(dist, dev.elapsed_ms(f0, f1))
    wall_ms = (time.perf_counter() - t0) * 1e3
    for g in e2e_graphs.values():
        g.close()
    if graph is not None:
        graph.close()
    if dp is not None:
        dp.close()

    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    out = {
        "metric": METRIC, "value": GLOBAL_BATCH * steps / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C4 MLP 1024-4096-4096-1024 where-ReLU mean-MSE SGD full training step",
                   "global_batch": GLOBAL_BATCH, "per_gpu_batch": local, "parallelism": f"dp{world}",
                   "l2": "working set per step (inputs+activations+grads >= 1.3 GB per GPU) exceeds "
                         "the 126 MB L2, no flush needed"},
        "loss": loss_value,
        "ms_per_step_by_rank": rank_ms,
        "e2e": {"value": GLOBAL_BATCH * steps / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int(feeder.bytes_per_step), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / steps, "wall_ms_per_step": wall_ms / steps},
        "gpu_launches": launches,
        "memory": mem,
        "clocks": clocks,
        "samples_per_s_per_sm_mhz": (GLOBAL_BATCH * steps / (ms * 1e-3)) / clocks["sm_mhz"] if clocks.get("sm_mhz") else None,
        "dp_parity": dp_parity,
        "step_mode": ("cuda_graph_replay (md.capture_graph of the whole step incl. the NCCL exchange)" if graph is not None
                      else "eager (one launch per op)" + (f"; graph capture failed: {graph_error}" if graph_error else "")),
        "roofline": tensor_roofline(
            gemm_tflops, peaks, kernel="mdb_gemm (matmul fwd + dW/dX gradient GEMMs): tc::gemm_3xtf32_pair_kernel",
            traffic=GEMM_TRAFFIC_BYTES_PER_LAUNCH if world == 1 else None,
            traffic_src=GEMM_TRAFFIC_SRC,
            algorithmic_bytes_per_launch=GEMM_ALGORITHMIC_BYTES_PER_LAUNCH if world == 1 else None,
            launches=int(gemm_n), avg_launch_ms=gemm_ms / gemm_n if gemm_n else None,
            share_of_step=(gemm_ms / prof_steps) / (ms / steps) if ms else None, gemm_paths=gemm_paths),
        "other_kernels": {
            "elementwise": {"ms_per_step": ew_ms / prof_steps, "calls_per_step": ew_n / prof_steps,
                            "algorithmic_GBps": ew_bytes / (ew_ms * 1e-3) / 1e9 if ew_ms else None,
                            "frac_of_hbm": ew_bytes / (ew_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if ew_ms else None},
            "reduce": {"ms_per_step": red_ms / prof_steps, "calls_per_step": red_n / prof_steps,
                       "algorithmic_GBps": red_bytes / (red_ms * 1e-3) / 1e9 if red_ms else None,
                       "frac_of_hbm": red_bytes / (red_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if red_ms else None},
        },
    }
    return out


def dp_parity_check(dev, dist, rank, world, dp, X, Y, params, local):
    """Evidence that the data-parallel path computes the right thing on hardware (the 1-GPU test
    lease cannot): ONE backward over the sharded global batch with the NCCL exchange, then rank 0
    alone recomputes the gradient of the same GLOBAL batch (all ranks' shards regenerated from their
    seeds) on its single GPU and compares.  max_rel = max over parameters of max|dp - single| /
    max|single|; summation order differs (per-shard means averaged by NCCL), so ~1e-6 is expected."""
    md = dev.md
    from minidiff_b200 import workloads as W

    loss = md.mean((W.mlp_forward(X, params) - Y) ** 2)
    loss.backward()
    dp.finish()
    dev.sync()
    dp_grads = [p.grad.as_numpy() for p in params]
    dp_loss = float(loss.item())
    losses = all_ranks(dist, dp_loss)
    for p in params:
        p.grad = None
    barrier(dist)
    out = None
    if rank == 0:
        shards = [W.mlp_data(local, DIMS[0], DIMS[-1], seed=1000 + 2 * r) for r in range(world)]
        Xg = md.Tensor(np.concatenate([s[0] for s in shards]))
        Yg = md.Tensor(np.concatenate([s[1] for s in shards]))
        del shards
        solo = [md.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]      # no grad hooks: no NCCL
        l1 = md.mean((W.mlp_forward(Xg, solo) - Yg) ** 2)
        l1.backward()
        worst, worst_rms = 0.0, 0.0
        for g, q in zip(dp_grads, solo):
            ref = q.grad.as_numpy().astype(np.float64)
            d = np.abs(g.astype(np.float64) - ref)
            worst = max(worst, float(d.max() / np.abs(ref).max()))
            worst_rms = max(worst_rms, float(np.sqrt((d ** 2).mean()) / np.sqrt((ref ** 2).mean())))
        single_loss = float(l1.item())
        out = {"max_rel": worst, "rms_rel": worst_rms, "ok": bool(worst < 1e-4),
               "loss_single_gpu": single_loss, "loss_mean_over_ranks": float(np.mean(losses)),
               "loss_rel_diff": abs(float(np.mean(losses)) - single_loss) / abs(single_loss),
               "what": f"averaged gradients of one DP backward (world {world}, {local} rows per rank) vs a "
                       f"single-GPU backward over the same {local * world}-row global batch on rank 0"}
        del Xg, Yg, solo, l1
    barrier(dist)
    return out


def bench_c2(dev, steps, warmup, peaks):
    md = dev.md
    from minidiff_b200 import workloads as W

    a_np, c_np = W.c2_inputs()
    a, c = md.Tensor(a_np, allow_grad=True), md.Tensor(c_np, allow_grad=True)
    warm_until_stable(dev, lambda: W.c2_step(a, c), warmup)
    time.sleep(0.3)          # let the power state settle after the GEMM-heavy headline loop
    # pass 1: whole-iteration device time, no per-launch profiler events on the stream (the event
    # pairs cost ~5 us per launch, 8 % of this 13-launch iteration)
    e0, e1 = dev.event(), dev.event()
    l0 = dev.launches()
    t0 = time.perf_counter()
    dev.record(e0)
    for _ in range(steps):
        loss = W.c2_step(a, c)
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    dev.record(e1)
    dev.sync()
    ms = dev.elapsed_ms(e0, e1) / steps
    launches = (dev.launches() - l0) / steps
    # pass 2: per-class split and the algorithmic bytes of OUR launches (profiler on)
    dev.prof(True)
    for _ in range(steps):
        W.c2_step(a, c)
    dev.sync()
    ew_ms, ew_n, ew_bytes = dev.prof_read(0)
    red_ms, red_n, red_bytes = dev.prof_read(1)
    dev.prof(False)
    kern_ms, kern_bytes = (ew_ms + red_ms) / steps, (ew_bytes + red_bytes) / steps
    return {
        "workload": "C2 sum(sin(a*c+a)**2).backward(), a:(8192,1) c:(1,8192) fp32 (2^26-element tensors)",
        "ms_per_iter": ms, "host_issue_ms_per_iter": host_ms, "launches_per_iter": launches,
        "loss": float(loss.item()),
        "reference_chain_GBps": W.C2_ALGORITHMIC_BYTES / (ms * 1e-3) / 1e9,
        "reference_chain_bytes": W.C2_ALGORITHMIC_BYTES,
        "note": "reference_chain_GBps = the 26E bytes the reference's 22-call chain moves (SURVEY 8d) "
                "/ our time; fused backward kernels move fewer bytes, so this can exceed the HBM peak. "
                "The roofline below is for OUR launches: their algorithmic bytes / the whole-iteration "
                "device time (CUDA events around the loop, launch gaps included).",
        "roofline": {"bound": "hbm", "achieved": kern_bytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": kern_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "traffic": C2_TRAFFIC_BYTES_PER_ITER, "traffic_src": C2_TRAFFIC_SRC,
                     "algorithmic_bytes_per_iter": kern_bytes,
                     "kernel_event_ms_per_iter": kern_ms,
                     "inputs": "8 tensors of 256 MiB per iteration > 126 MB L2 (no flush needed)"},
    }


def bench_c3(dev, steps, warmup, peaks, n=8192):
    md = dev.md
    from minidiff_b200 import workloads as W

    A_np, B_np = W.c3_inputs(n)
    A, B = md.Tensor(A_np, allow_grad=True), md.Tensor(B_np, allow_grad=True)
    warm_until_stable(dev, lambda: W.c3_step(A, B), warmup)
    e0, e1 = dev.event(), dev.event()
    dev.prof(True)
    dev.record(e0)
    for _ in range(steps):
        W.c3_step(A, B)
    dev.record(e1)
    dev.sync()
    ms = dev.elapsed_ms(e0, e1) / steps
    g_ms, g_n, g_fl = dev.prof_read(2)
    dev.prof(False)
    tf = g_fl / (g_ms * 1e-3) / 1e12
    return {"workload": f"C3 C=A@B; C.backward() {n}^3 fp32 (NN fwd, NT dA, TN dB)",
            "ms_per_iter": ms, "TFLOPs_fp32_equiv": W.c3_flops(n) / (ms * 1e-3) / 1e12,
            "gemm_paths": dev.gemm_paths(),
            "roofline": tensor_roofline(tf, peaks, traffic=C3_TRAFFIC_BYTES_PER_LAUNCH, traffic_src=C3_TRAFFIC_SRC,
                                        algorithmic_bytes_per_launch=3.0 * 4.0 * n * n, avg_launch_ms=g_ms / g_n)}


def bench_c1(dev, iters=200):
    """BASELINE config 1 (latency-bound, no roofline): README example on 2x4 tensors, first- and
    second-order backward; eager (one Python call + launch per op) vs one CUDA-graph replay."""
    md = dev.md
    x = md.Tensor([[0, 2, -2, 1], [-1, -1, -2, -2]], allow_grad=True, dtype=md.float32)
    y = md.Tensor([[2, 3, 4, 5], [0, -1, -3, 2]], allow_grad=True, dtype=md.float32)

    def step():
        f = 2 * y * md.sin(x) - x ** 2
        f.backward(allow_higher_order=True)
        x.grad.backward()
        return f

    for _ in range(5):
        step()
    dev.sync()
    l0 = dev.launches()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dev.sync()
    eager_us = (time.perf_counter() - t0) * 1e6 / iters
    launches = (dev.launches() - l0) / iters
    g = md.capture_graph(step)
    g.replay()
    dev.sync()
    e0, e1 = dev.event(), dev.event()
    t0 = time.perf_counter()
    dev.record(e0)
    for _ in range(iters):
        g.replay()
    dev.record(e1)
    dev.sync()
    graph_wall_us = (time.perf_counter() - t0) * 1e6 / iters
    graph_dev_us = dev.elapsed_ms(e0, e1) * 1e3 / iters
    d2 = x.grad.as_numpy().ravel().tolist()
    g.close()
    return {"workload": "C1 f = 2*y*sin(x) - x**2 on 2x4 fp32, backward(allow_higher_order) + x.grad.backward()",
            "launches_per_iter": launches, "eager_us_per_iter": eager_us,
            "graph_replay_us_per_iter": graph_wall_us, "graph_replay_device_us_per_iter": graph_dev_us,
            "d2f_dx2": d2}


def bench_c4_reference_engine(dev, steps, warmup):
    """The pure drop-in number: the UNMODIFIED reference engine (baseline/_ref, its own Tensor /
    OpNode / ops/definitions.py, out-of-place gradient accumulation, unfused backward) with only
    `--backend minidiff_b200.plugin` switched, on the same C4 training step.  None if the offline
    install of the reference is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "minidiff")):
        return None
    for p in (os.path.join(ROOT, "oracle", "_stubs"), ref_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    saved = sys.argv
    sys.argv = [saved[0], "--backend", "minidiff_b200.plugin"]
    try:
        import minidiff as ref
    finally:
        sys.argv = saved
    import minidiff_b200.plugin as plugin
    from minidiff_b200 import workloads as W

    plugin.assert_live(ref)
    X_np, Y_np = W.mlp_data(GLOBAL_BATCH, DIMS[0], DIMS[-1], seed=1000)
    params = [ref.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]
    X, Y = ref.Tensor(X_np), ref.Tensor(Y_np)

    def step():
        h = X
        for l in range(3):
            h = h @ params[2 * l] + params[2 * l + 1]
            if l < 2:
                h = ref.where(h > 0, h, 0)
        loss = ref.mean((h - Y) ** 2)
        loss.backward()
        with ref.no_grad():
            for p in params:
                p -= LR * p.grad
        return loss

    import gc

    loss = warm_until_stable(dev, step, warmup)
    # the reference engine frees part of each step's graph through Python's CYCLE collector, so the
    # moment device blocks return to the caching allocator is not deterministic; a timed pass that
    # hit a cudaMalloc (a device-wide synchronisation) is repeated, at most three times
    for attempt in range(3):
        gc.collect()
        dev.sync()
        e0, e1 = dev.event(), dev.event()
        l0, a0 = dev.launches(), dev.mem()["device_allocs"]
        dev.record(e0)
        for _ in range(steps):
            loss = step()
        dev.record(e1)
        dev.sync()
        ms = dev.elapsed_ms(e0, e1) / steps
        allocs = dev.mem()["device_allocs"] - a0
        if allocs == 0:
            break
    return {"workload": "C4 training step, UNMODIFIED reference engine + --backend minidiff_b200.plugin",
            "ms_per_step": ms, "samples_per_s": GLOBAL_BATCH / (ms * 1e-3),
            "launches_per_step": (dev.launches() - l0) / steps, "loss": float(loss.item()),
            "device_allocs_in_timed_region": allocs, "timed_passes": attempt + 1}


def bench_c5(dev, steps, warmup, peaks, batch=8192):
    """BASELINE config 5: Hessian-vector product through a second-order graph of device ops."""
    md = dev.md
    from minidiff_b200 import workloads as W

    X_np, Y_np = W.mlp_data(batch, DIMS[0], DIMS[-1], seed=40)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]
    vs = [md.Tensor(np.random.default_rng(50 + i).standard_normal(p.shape).astype(np.float32))
          for i, p in enumerate(params)]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    warm_until_stable(dev, lambda: W.hvp(X, Y, params, vs), warmup)
    e0, e1 = dev.event(), dev.event()
    dev.prof(True)
    l0 = dev.launches()
    allocs0 = dev.mem()["device_allocs"]
    dev.record(e0)
    for _ in range(steps):
        hv = W.hvp(X, Y, params, vs)
    dev.record(e1)
    dev.sync()
    ms = dev.elapsed_ms(e0, e1) / steps
    g_ms, g_n, g_fl = dev.prof_read(2)
    dev.prof(False)
    tf = g_fl / (g_ms * 1e-3) / 1e12
    return {"workload": f"C5 Hessian-vector product, same MLP, batch {batch}, allow_higher_order backward "
                        "then backward of sum(grad*v)",
            "ms_per_iter": ms, "launches_per_iter": (dev.launches() - l0) / steps,
            "gemm_launches_per_iter": g_n / steps, "gemm_flops_per_iter": g_fl / steps,
            "device_allocs_in_timed_region": dev.mem()["device_allocs"] - allocs0,
            "hv_norm": float(md.sum(hv[0] * hv[0]).item()) ** 0.5,
            "gemm_paths": dev.gemm_paths(),
            "roofline": tensor_roofline(tf, peaks, traffic=None, gemm_share_of_iter=(g_ms / steps) / ms)}


def bench_c4_fused(dev, steps, warmup, peaks, fused=True, split="3xtf32"):
    """fused=True: the C4 training step written with the stateful fused ops: md.linear_relu(X, W, b)
    = one tcgen05 GEMM with bias + ReLU in its epilogue, md.linear for the output layer; backward
    masks the upstream gradient once per layer (or inside the dX GEMM epilogue of the layer above).
    fused=False: the headline step (workloads.mlp_train_step).  split: operand split of the tensor-core GEMM
    (backend.set_matmul_split): "3xtf32" is the default everywhere else in this file; "fast" = one TF32 MMA + two
    BF16 cross-term MMAs per product, opt-in because its error sits at the edge of the GEMM tolerance."""
    md = dev.md
    from minidiff_b200 import workloads as W

    X_np, Y_np = W.mlp_data(GLOBAL_BATCH, DIMS[0], DIMS[-1], seed=1000)
    params = [md.Tensor(p, allow_grad=True) for p in W.mlp_params(DIMS)]
    X, Y = md.Tensor(X_np), md.Tensor(Y_np)
    del X_np, Y_np

    def step():
        if not fused:
            return W.mlp_train_step(X, Y, params, LR, None)
        h = md.linear_relu(X, params[0], params[1])
        h = md.linear_relu(h, params[2], params[3])
        loss = md.mean((md.linear(h, params[4], params[5]) - Y) ** 2)
        loss.backward()
        with md.no_grad():
            for p in params:
                p -= LR * p.grad
        return loss

    md.backend.set_matmul_split(split)
    try:
        loss = warm_until_stable(dev, step, warmup)
        e0, e1 = dev.event(), dev.event()
        dev.prof(True)
        dev.gemm_paths(reset=True)
        l0 = dev.launches()
        dev.record(e0)
        for _ in range(steps):
            loss = step()
        dev.record(e1)
        dev.sync()
    finally:
        md.backend.set_matmul_split("3xtf32")
    ms = dev.elapsed_ms(e0, e1) / steps
    g_ms, g_n, g_fl = dev.prof_read(2)
    ew_ms, ew_n, ew_b = dev.prof_read(0)
    red_ms, red_n, red_b = dev.prof_read(1)
    dev.prof(False)
    tf = g_fl / (g_ms * 1e-3) / 1e12
    roof = tensor_roofline(tf, peaks, launches=int(g_n), share_of_step=(g_ms / steps) / ms)
    name = "C4 training step written with md.linear_relu / md.linear (fused stateful ops)" if fused else \
        "C4 training step (same code as the headline)"
    if split == "fast":
        name += "; GEMM operand split 'fast' (backend.set_matmul_split)"
        roof["pipe_frac"] = 2.0 * roof["frac"]
        roof["note"] = ("fast split: one TF32 MMA + two BF16 MMAs (half the tensor-pipe time each) per fp32 product = 2 "
                        "TF32-equivalents, so pipe_frac = 2*frac; rms error 1.4e-6 of the result's rms instead of 0.5e-6 "
                        "(tests/test_gpu_gemm.py::test_operand_splits_accuracy_against_float64)")
    return {"workload": name, "gemm_split": split,
            "ms_per_step": ms, "samples_per_s": GLOBAL_BATCH / (ms * 1e-3),
            "launches_per_step": (dev.launches() - l0) / steps, "loss": float(loss.item()),
            "gemm_paths": dev.gemm_paths(),
            "elementwise_ms_per_step": ew_ms / steps, "reduce_ms_per_step": red_ms / steps,
            "roofline": roof}


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (baseline/_ref) on its NumPy backend, all host threads
# ------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1, which would pin OpenBLAS (the reference's matmul) to one
    thread; the CPU arm is meant to use every host core.  Returns the thread count in effect."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=n)
        got = [i.get("num_threads") for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(got) if got else n
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


# synthetic inputs of the BASELINE configs (pure NumPy; identical to minidiff_b200/workloads.py, repeated
# here so that the reference arm imports nothing of this repo's engine)
def gen_c2_inputs(n=8192, m=8192, seed=1234):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, 1)).astype(np.float32), rng.standard_normal((1, m)).astype(np.float32))


def gen_c3_inputs(n=8192):
    return (np.random.default_rng(1234).standard_normal((n, n), dtype=np.float32),
            np.random.default_rng(1235).standard_normal((n, n), dtype=np.float32))


def gen_mlp_params(dims=DIMS, seed0=2):
    ps, sd = [], seed0
    for fi, fo in zip(dims[:-1], dims[1:]):
        ps.append((np.random.default_rng(sd).standard_normal((fi, fo)) / np.sqrt(fi)).astype(np.float32))
        ps.append((np.random.default_rng(sd + 1).standard_normal((fo,)) / np.sqrt(fo)).astype(np.float32))
        sd += 2
    return ps


def gen_mlp_data(batch, d_in, d_out, seed=0):
    return (np.random.default_rng(seed).standard_normal((batch, d_in), dtype=np.float32),
            np.random.default_rng(seed + 1).standard_normal((batch, d_out), dtype=np.float32))


def import_reference_numpy():
    """The unmodified reference package from baseline/_ref (offline pip install of /root/reference
    done by __graft_entry__.build(); shipped to the GPU box) on ITS OWN NumPy backend: argv is trimmed
    because the reference parses sys.argv at import (backend/__init__.py:13-16), graphviz is stubbed
    (only utils.py drawing uses it).  None if the install is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "minidiff")):
        return None
    for p_ in (os.path.join(ROOT, "oracle", "_stubs"), ref_dir):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    saved = sys.argv
    sys.argv = [saved[0]]
    try:
        import minidiff as ref
    finally:
        sys.argv = saved
    import minidiff.backend as live

    if live.tensor_class is not np.ndarray:
        raise RuntimeError("the reference did not select its NumPy backend")
    return ref


class PortEngine:
    """fallback when baseline/_ref is absent: the oracle port of the same algorithm (kind 'port')"""

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import np_minidiff as orc

        self.orc = orc


def ref_workload(ref, name, rows=None):
    """(step function, units per step, unit, description) of BASELINE config `name` written against
    the reference's public API -- the same user code as minidiff_b200/workloads.py."""
    c2_inputs, c3_inputs, mlp_data, mlp_params = gen_c2_inputs, gen_c3_inputs, gen_mlp_data, gen_mlp_params
    if name == "c4":
        B = rows or GLOBAL_BATCH
        X_np, Y_np = mlp_data(B, DIMS[0], DIMS[-1], seed=1000)
        params = [ref.Tensor(p_, allow_grad=True) for p_ in mlp_params(DIMS)]
        X, Y = ref.Tensor(X_np), ref.Tensor(Y_np)

        def step():
            h = X
            for l in range(3):
                h = h @ params[2 * l] + params[2 * l + 1]
                if l < 2:
                    h = ref.where(h > 0, h, 0)
            loss = ref.mean((h - Y) ** 2)
            loss.backward()
            with ref.no_grad():
                for p_ in params:
                    p_ -= LR * p_.grad
            return loss

        return step, B, "samples/s", f"one full C4 training step on {B} rows"
    if name == "c2":
        a_np, c_np = c2_inputs()
        a, c = ref.Tensor(a_np, allow_grad=True), ref.Tensor(c_np, allow_grad=True)

        def step():
            loss = ref.sum(ref.sin(a * c + a) ** 2)
            loss.backward()
            return loss

        return step, 26.0 * (1 << 26) * 4 / 1e9, "GB/s", "one C2 iteration (26E reference-chain bytes)"
    if name == "c3":
        A_np, B_np = c3_inputs(8192)
        A, Bm = ref.Tensor(A_np, allow_grad=True), ref.Tensor(B_np, allow_grad=True)

        def step():
            Cm = A @ Bm
            Cm.backward()
            return Cm

        return step, 3 * 2.0 * 8192 ** 3 / 1e12, "TFLOP/s", "one C3 iteration (8192^3 fwd + 2 gradient GEMMs)"
    if name == "c5":
        B = rows or 8192
        X_np, Y_np = mlp_data(B, DIMS[0], DIMS[-1], seed=40)
        params = [ref.Tensor(p_, allow_grad=True) for p_ in mlp_params(DIMS)]
        vs = [ref.Tensor(np.random.default_rng(50 + i).standard_normal(p_.shape).astype(np.float32))
              for i, p_ in enumerate(params)]
        X, Y = ref.Tensor(X_np), ref.Tensor(Y_np)

        def step():
            h = X
            for l in range(3):
                h = h @ params[2 * l] + params[2 * l + 1]
                if l < 2:
                    h = ref.where(h > 0, h, 0)
            L = ((h - Y) ** 2) / float(h.size)
            L.backward(allow_higher_order=True)
            s_ = None
            for p_, v in zip(params, vs):
                t = ref.sum(p_.grad * v)
                s_ = t if s_ is None else s_ + t
            s_.backward()
            return s_

        flops = 2.0 * B * (71303168 + 134217728)
        return step, flops / 1e12, "TFLOP/s", f"one C5 Hessian-vector product, batch {B} (22 GEMMs)"
    raise ValueError(name)


def reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the workload, timed on this box's
    host cores.  Rank 0 only.  C4 runs the FULL 65536-row batch per step (same config as the B200
    arm); only if one step is so slow that K+W steps would not end within --cpu-budget-s is the
    per-step sample reduced (and the line says so)."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    ref = import_reference_numpy()
    kind, src = "reference", "unmodified reference from baseline/_ref on its NumPy backend"
    if ref is None:
        return port_arm(args, cores)
    name = args.workload
    rows = args.sample_rows or None
    step, units, unit, what = ref_workload(ref, name, rows)
    t0 = time.perf_counter()
    step()                                              # first warm-up step, also the cost probe
    first = time.perf_counter() - t0
    todo = args.warmup - 1 + args.steps
    if name in ("c4", "c5") and rows is None and first * todo > args.cpu_budget_s:
        full = GLOBAL_BATCH if name == "c4" else 8192
        rows = max(2048, int(full * args.cpu_budget_s / (first * todo)) // 2048 * 2048)
        del step
        step, units, unit, what = ref_workload(ref, name, rows)
        step()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    v = units * len(times) / total
    full_cfg = (name != "c4" or rows in (None, GLOBAL_BATCH)) and (name != "c5" or rows in (None, 8192))
    line = {
        "impl": "reference", "metric": METRIC if name == "c4" else f"{name}_cpu_reference", "value": v, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
        "best_ms_per_step": min(times) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 MLP 1024-4096-4096-1024 where-ReLU mean-MSE SGD full training step" if name == "c4" else what,
                   "global_batch": GLOBAL_BATCH if name == "c4" else None,
                   "rows_per_step": rows or (GLOBAL_BATCH if name == "c4" else None),
                   "full_batch": full_cfg, "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": kind,
                         "sample": f"{what}; {len(times)} timed steps after {args.warmup} warm-up; {src} "
                                   f"(NumPy {np.__version__} / OpenBLAS: matmul on all {cores} host threads, "
                                   "elementwise + reductions single-threaded as NumPy runs them)"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "result_check": float(np.asarray(out.as_numpy() if hasattr(out, "as_numpy") else out).ravel()[0]),
    }
    print(json.dumps(line), flush=True)


def port_arm(args, cores):
    """baseline/_ref missing: time the oracle port (same algorithm on NumPy) instead, C4 only."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_minidiff as orc

    rows = args.sample_rows or 8192
    X, Y = orc.mlp_data(rows, DIMS[0], DIMS[-1])
    ps = orc.mlp_params(DIMS)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.config4_step(X, Y, ps)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    v = rows * len(times) / sum(times)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(times) / len(times) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 MLP 1024-4096-4096-1024 where-ReLU mean-MSE SGD full training step",
                   "global_batch": GLOBAL_BATCH, "rows_per_step": rows, "full_batch": rows == GLOBAL_BATCH,
                   "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"baseline/_ref absent: oracle port, one training step on {rows} rows"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def cpu_baseline_leg(workload, steps=2, warmup=1, timeout=600):
    """Run `bench.py --impl reference --workload X` in a SUBPROCESS (this process has already bound the
    reference package to the B200 plugin for the drop-in leg; one backend per process) and return its
    cpu_baseline object."""
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
           "--steps", str(steps), "--warmup", str(warmup)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                cb = d["cpu_baseline"]
                cb["ms_per_step"] = d["ms_per_step"]
                cb["full_batch"] = d["config"].get("full_batch")
                return cb
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as exc:
        return {"error": repr(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4", "c5"],
                    help="--impl reference only: which BASELINE config the CPU reference runs")
    ap.add_argument("--sample-rows", type=int, default=0, help="--impl reference: rows per step (0 = full batch)")
    ap.add_argument("--cpu-budget-s", type=float, default=900.0,
                    help="--impl reference: shrink the per-step sample if K+W full-batch steps would exceed this")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="run the timed C4 step as one CUDA-graph replay (auto: multi-GPU runs only)")
    ap.add_argument("--skip-extras", action="store_true", help="skip the C1/C2/C3/C5 single-GPU benchmarks")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return reference_arm(args, rank)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    warmup = max(args.warmup, 3)
    peaks = load_peaks()
    sys.argv = sys.argv[:1]
    dev = Dev()
    dist = dist_setup(world)
    use_graph = args.graph == "on" or (args.graph == "auto" and world > 1)
    line = bench_mlp(dev, dist, rank, world, args.steps, warmup, peaks, use_graph)
    line["warmup"] = warmup
    if rank == 0 and world == 1:
        if not args.skip_extras:
            # cuBLAS TF32 on this box, a SIDE NOTE (taken after the headline loop so it cannot pre-heat it);
            # every frac in this file uses MEASURED_PEAKS.json (tensor_roofline)
            line["roofline"]["tf32_cublas_side_note"] = measure_tf32_peak()
            time.sleep(1.0)
            line["fwd_bwd"] = {"c1_readme": bench_c1(dev),
                               "c2_broadcast_chain": bench_c2(dev, max(args.steps, 10), warmup, peaks),
                               "c3_matmul": bench_c3(dev, max(2, min(args.steps, 5)), warmup, peaks),
                               "c5_hvp": bench_c5(dev, max(args.steps, 10), warmup, peaks),
                               "c4_fused": bench_c4_fused(dev, max(5, min(args.steps, 10)), warmup, peaks),
                               "c4_fast_split": bench_c4_fused(dev, max(5, min(args.steps, 10)), warmup, peaks,
                                                               fused=False, split="fast"),
                               "c4_fused_fast_split": bench_c4_fused(dev, max(5, min(args.steps, 10)), warmup, peaks,
                                                                     fused=True, split="fast"),
                               "c4_reference_engine_dropin": bench_c4_reference_engine(dev, 5, warmup)}
        if not args.skip_cpu:
            # the unmodified reference on its NumPy backend, on this box's host cores, same workloads
            dev.md.backend.functions.synchronize()
            line["cpu_baseline"] = cpu_baseline_leg("c4")
            if not args.skip_extras:
                for key, wl in (("c2_broadcast_chain", "c2"), ("c3_matmul", "c3"), ("c5_hvp", "c5")):
                    line["fwd_bwd"][key]["cpu_baseline"] = cpu_baseline_leg(wl)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
