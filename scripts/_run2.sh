cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused_ops.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests2.log
cat gpurun_out/r2_tests2.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_bench_a.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
fb=d.pop('fwd_bwd',{})
print(json.dumps(d,indent=1)[:5000])
for k,v in fb.items(): print(k, json.dumps(v)[:1500])
P
