cd $GRAFT_REPO_ROOT
for mode in on off; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 30 --warmup 5 --graph $mode > gpurun_out/r2_bench_n8_$mode.json 2> gpurun_out/r2_bench_n8_$mode.err
tail -2 gpurun_out/r2_bench_n8_$mode.err
python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_n8_$mode.json'))
for k in ('value','ms_per_step','step_mode','gpu_launches','loss','clocks','ms_per_step_by_rank'): print(k, d.get(k))
print(d['e2e'], d['dp_parity'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline']['gemm_paths'], d['other_kernels'])
P
done
