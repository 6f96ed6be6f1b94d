cd $GRAFT_REPO_ROOT
for mode in on off; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 --graph $mode > gpurun_out/r2_bench_n2_$mode.json 2> gpurun_out/r2_bench_n2_$mode.err
tail -3 gpurun_out/r2_bench_n2_$mode.err
python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_n2_$mode.json'))
for k in ('value','ms_per_step','step_mode','gpu_launches','loss'): print(k, d.get(k))
print(d['e2e'], d['dp_parity']['max_rel'], d['roofline']['achieved'], d['roofline']['share_of_step'])
P
done
