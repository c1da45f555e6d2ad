CMD="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths' -s 1 -c 1 -o gpurun_out/r1p_paths -f $CMD > gpurun_out/ncu_p.log 2>&1
tail -n 3 gpurun_out/ncu_p.log
