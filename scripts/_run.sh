set -x
timeout 300 python scripts/gemm_split_check.py > gpurun_out/split_check4.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu > gpurun_out/r2_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu > gpurun_out/r2_ncu_list.log 2>&1
timeout 300 python scripts/profile_step.py mlp > gpurun_out/r2_plain2.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_3xtf32_pair -s 8 -c 8 -o /tmp/r02_mlp_gemm -f python scripts/profile_step.py mlp > gpurun_out/r2_ncu_full1.log 2>&1
python scripts/ncu_summary.py /tmp/r02_mlp_gemm.ncu-rep > gpurun_out/r02_ncu_mlp_step_gemm.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_3xtf32_pair -s 3 -c 3 -o gpurun_out/r02_c3_gemm -f python scripts/profile_step.py c3 > gpurun_out/r2_ncu_full2.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_c3_gemm.ncu-rep > gpurun_out/r02_ncu_c3_gemm.md
MDB_GEMM_SPLIT=fast timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_3xtf32_pair -s 3 -c 3 -o gpurun_out/r02_c3_gemm_fast -f python scripts/profile_step.py c3 > gpurun_out/r2_ncu_full3.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_c3_gemm_fast.ncu-rep > gpurun_out/r02_ncu_c3_gemm_fast_split.md
du -sh gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
