"""C2 chain: whole-iteration device time with and without the per-launch profiler events, and host
issue time (is the chain launch-bound?).   python scripts/c2_timing.py"""
import sys, time
sys.path.insert(0, ".")
sys.argv = sys.argv[:1]
import minidiff_b200 as md
from minidiff_b200 import workloads as W
from bench import Dev
dev = Dev()
a_np, c_np = W.c2_inputs()
a, c = md.Tensor(a_np, allow_grad=True), md.Tensor(c_np, allow_grad=True)
for _ in range(5):
    W.c2_step(a, c)
dev.sync()
for prof in (False, True, False):
    dev.prof(prof)
    e0, e1 = dev.event(), dev.event()
    n = 30
    t0 = time.perf_counter()
    dev.record(e0)
    for _ in range(n):
        W.c2_step(a, c)
    t_issue = time.perf_counter() - t0
    dev.record(e1)
    dev.sync()
    t_wall = time.perf_counter() - t0
    extra = ""
    if prof:
        ew = dev.prof_read(0); rd = dev.prof_read(1)
        extra = f" | event-pair sums: ew {ew[0]/n:.3f} ms ({ew[1]/n:.0f} calls) red {rd[0]/n:.3f} ms ({rd[1]/n:.0f} calls)"
    dev.prof(False)
    print(f"prof={prof}: device {dev.elapsed_ms(e0,e1)/n:.3f} ms/iter, host issue {t_issue/n*1e3:.3f} ms/iter, wall {t_wall/n*1e3:.3f}{extra}", flush=True)
