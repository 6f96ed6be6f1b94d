python - <<'P'
import os, sys
sys.path.insert(0, ".")
os.environ["SWEEP"] = "split"
src = open("scripts/gemm_sweep.py").read().replace('("hybrid c1", dict(split=1, chunk=1))', '("hybrid c8", dict(split=1, chunk=8)), ("hybrid c16", dict(split=1, chunk=16))').replace('("3xtf32 c2", dict(split=0, chunk=2)), ', '').replace('("hybrid c2", dict(split=1, chunk=2)),', '')
sys.argv = ["gemm_sweep.py", "65536", "5", "fwd2,dW2,c3"]
exec(compile(src, "gemm_sweep.py", "exec"))
P
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "exact or split or compensation" 2>&1 | tail -2
