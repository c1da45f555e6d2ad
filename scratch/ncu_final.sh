CMD="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_f3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_sweep|k_march|k_path|k_terrain|k_ray|k_column|k_prepare' -s 11 -c 24 --csv --log-file gpurun_out/r01_c5_launches.csv $CMD > gpurun_out/ncu_f3l.log 2>&1
$CMD > gpurun_out/plain_f4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths|k_terrain_profile|k_sweep' -s 4 -c 4 -o gpurun_out/r1f_prof -f $CMD > gpurun_out/ncu_f3.log 2>&1
tail -n 2 gpurun_out/ncu_f3.log
