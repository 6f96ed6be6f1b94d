cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_dp.py -q -m gpu 2>&1 | tail -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
for k in ('value','ms_per_step','e2e','dp_parity','clocks','gpu_launches','ms_per_step_by_rank'): print(k, d.get(k))
print(d['roofline']['achieved'], d['roofline']['gemm_paths'], d['other_kernels'])
P
