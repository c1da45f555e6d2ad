CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_sweep|k_march|k_path|k_terrain|k_ray|k_column|k_prepare' -s 12 -c 14 --csv --log-file gpurun_out/r1f_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
