cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_gemm_tests.log
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "baseline or c5 or c4" 2>&1 | tail -15 > gpurun_out/r2_engine_tests.log
timeout 900 python scripts/gemm_sweep.py 65536 7 > gpurun_out/r2_sweep_65536.log 2>&1
cat gpurun_out/r2_gemm_tests.log gpurun_out/r2_engine_tests.log gpurun_out/r2_sweep_65536.log
