timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputests.txt 2>&1; tail -3 gpurun_out/gputests.txt
timeout 300 python scripts/microbench_ew.py 20 > gpurun_out/microbench_ew.txt 2>&1; cat gpurun_out/microbench_ew.txt
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('other_kernels'))"
