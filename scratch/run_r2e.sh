# round 2, call E: sweep v3 (incremental pointers, next-group prefetch, four-window fast-forward) + dumper tests
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2e_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for mb in 10 8 12; do ATMRT_SWEEP_MB=$mb $B > gpurun_out/r2e_new_mb$mb.json 2>> gpurun_out/r2e_var.err; done
python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_c2.json 2>> gpurun_out/r2e_var.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_c4.json 2>> gpurun_out/r2e_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sweep_bits|k_hit_normals|k_shade_tiles' -s 3 -c 3 -o gpurun_out/r2e_prof -f $C > gpurun_out/r2e_ncu.log 2>&1
tail -n 3 gpurun_out/r2e_ncu.log
