CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sweep' -s 2 -c 2 -o gpurun_out/r1g_sweep -f $CMD > gpurun_out/ncu_c.log 2>&1
tail -n 2 gpurun_out/ncu_c.log
