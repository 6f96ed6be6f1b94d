# Where does the data-parallel step lose time?  Per-rank batch of the 8-GPU run (8192 rows) on 2 GPUs
# (global batch 16384), varying how many SMs NCCL may use and how many the GEMM leaves free.
cd ${GRAFT_REPO_ROOT:-.}
run() {  # label, extra env...
  label=$1; shift
  env "$@" MDB_BENCH_GLOBAL_BATCH=16384 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 40 --warmup 5 --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['step_mode'][:12], 'gemm TF %.1f' % d['roofline']['achieved'], 'parity', d['dp_parity']['max_rel'])"
}
run baseline A=1
run baseline2 A=1
run nchan2 NCCL_MAX_NCHANNELS=2
run nchan4 NCCL_MAX_NCHANNELS=4
run nchan8 NCCL_MAX_NCHANNELS=8
run maxctas4 NCCL_MAX_CTAS=4
run gemm72 MDB_BENCH_GEMM_MAX_CLUSTERS=72
run gemm70 MDB_BENCH_GEMM_MAX_CLUSTERS=70
run gemm70_nchan4 MDB_BENCH_GEMM_MAX_CLUSTERS=70 NCCL_MAX_NCHANNELS=4
run gemm66_nchan8 MDB_BENCH_GEMM_MAX_CLUSTERS=66 NCCL_MAX_NCHANNELS=8
