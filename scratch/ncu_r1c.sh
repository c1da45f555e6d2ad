set -x
CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_r1c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths|k_march|k_terrain_profile' -s 3 -c 3 -o gpurun_out/r1c_prof -f $CMD > gpurun_out/ncu_r1c.log 2>&1
tail -n 5 gpurun_out/ncu_r1c.log
