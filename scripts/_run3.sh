cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused_ops.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests3.log
cat gpurun_out/r2_tests3.log
