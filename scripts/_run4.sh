cd $GRAFT_REPO_ROOT
python scripts/gemm_traffic.py 65536 matrix fwd2,c3 > gpurun_out/r2_traffic2_launches.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gemm_3xtf32_pair --csv --log-file gpurun_out/r2_traffic2_ncu.csv python scripts/gemm_traffic.py 65536 matrix fwd2,c3 > gpurun_out/r2_traffic2_launches.log 2>&1
python scripts/ncu_traffic_table.py gpurun_out/r2_traffic2_launches.log gpurun_out/r2_traffic2_ncu.csv > gpurun_out/r2_traffic2_table.md
cat gpurun_out/r2_traffic2_table.md
