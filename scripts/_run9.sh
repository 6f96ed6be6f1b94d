cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu > gpurun_out/r2_ncu_launch.log 2>&1
python scripts/profile_step.py mlp > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_3xtf32_pair -s 8 -c 8 -o gpurun_out/r02_mlp_gemm python scripts/profile_step.py mlp > gpurun_out/r2_ncu_full.log 2>&1
python scripts/profile_step.py c3 > gpurun_out/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_3xtf32_pair -s 3 -c 3 -o gpurun_out/r02_c3_gemm python scripts/profile_step.py c3 > gpurun_out/r2_ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_ncu_full.log
