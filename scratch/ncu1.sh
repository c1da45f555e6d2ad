set -x
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1a_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths|k_march|k_terrain_profile' -c 3 -o gpurun_out/r1a_prof -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
ls -la gpurun_out
