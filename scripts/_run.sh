timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputests.txt 2>&1; tail -8 gpurun_out/gputests.txt
