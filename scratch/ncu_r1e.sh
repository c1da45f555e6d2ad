CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_r1e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1e_launches.csv $CMD > gpurun_out/ncu_r1e_l.log 2>&1
$CMD > gpurun_out/plain_r1e2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths|k_march|k_terrain_profile' -s 3 -c 3 -o gpurun_out/r1e_prof -f $CMD > gpurun_out/ncu_r1e.log 2>&1
tail -n 3 gpurun_out/ncu_r1e.log
