# round 2, call G: software-pipelined ray-path stage; group API compiled in
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2g_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r2g_n1.json 2>> gpurun_out/r2g_var.err
$B --emulate-ranks 8 > gpurun_out/r2g_e8.json 2>> gpurun_out/r2g_var.err
$B --emulate-ranks 2 > gpurun_out/r2g_e2.json 2>> gpurun_out/r2g_var.err
python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2g_c2.json 2>> gpurun_out/r2g_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths_macro|k_shade_tiles|k_hit_normals' -s 1 -c 3 -o gpurun_out/r2g_prof -f $C > gpurun_out/r2g_ncu.log 2>&1
tail -n 3 gpurun_out/r2g_ncu.log
