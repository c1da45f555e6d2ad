# round 2, call W: ncu --set full of the crossing march on c4
C="python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python profiles/source_sha.py > gpurun_out/r2w_sha.txt
$C > gpurun_out/r2w_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_cross_march|k_thresholds' -s 0 -c 2 -o gpurun_out/r2w_prof -f $C > gpurun_out/r2w_ncu.log 2>&1
tail -n 2 gpurun_out/r2w_ncu.log
