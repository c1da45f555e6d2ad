CMD="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march' -s 1 -c 1 -o gpurun_out/r1_c4 -f $CMD > gpurun_out/ncu_c4.log 2>&1
tail -n 2 gpurun_out/ncu_c4.log
