CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_q.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths|k_sweep|k_terrain_profile' -s 4 -c 4 -o gpurun_out/r1q_prof -f $CMD > gpurun_out/ncu_q.log 2>&1
tail -n 3 gpurun_out/ncu_q.log
