timeout 900 python -m pytest tests/test_dp.py -m gpu -q -rs > gpurun_out/r02_dp_tests_2gpu.txt 2>&1; tail -5 gpurun_out/r02_dp_tests_2gpu.txt
