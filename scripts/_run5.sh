cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_device_helpers.py -q -m gpu 2>&1 | tail -80 > gpurun_out/r2_tests5.log
cat gpurun_out/r2_tests5.log
