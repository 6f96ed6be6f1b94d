# N = 1, 2, 4, 8 back to back on ONE box (what the driver's scaling run does), headline legs only
cd ${GRAFT_REPO_ROOT:-.}
for n in 1 2 4 8; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py"; fi
  timeout 600 $cmd --gpus $n --steps 30 --warmup 5 --skip-cpu --skip-extras 2>/dev/null > gpurun_out/r02_scale_n$n.json
  python - <<P
import json
d=json.load(open("gpurun_out/r02_scale_n$n.json"))
print("N=$n ms/step %.3f value %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d["step_mode"][:10], "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "gemm TF %.1f" % d["roofline"]["achieved"], "parity", (d.get("dp_parity") or {}).get("max_rel"))
P
done
