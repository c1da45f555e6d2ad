# round 2, call 3X: validation of the current tree -- tests, default bench, reference arm, emulated rank of 8, c4, c2; ncu capture of the c5 stage kernels
python profiles/source_sha.py > gpurun_out/r3x_sha.txt
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r3x_tests.log
python bench.py > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3x_ref.json 2> gpurun_out/r3x_ref.err; echo "ref rc $?"
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --workload c5 --emulate-ranks 8 > gpurun_out/r3x_e8.json 2>> gpurun_out/r3x_var.err
$B --workload c4 > gpurun_out/r3x_c4.json 2>> gpurun_out/r3x_var.err
$B --workload c2 > gpurun_out/r3x_c2.json 2>> gpurun_out/r3x_var.err
$B --workload c1 > gpurun_out/r3x_c1.json 2>> gpurun_out/r3x_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r3x_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths_macro|k_terrain_profile|k_sweep_bits|k_hit_normals|k_shade_tiles' -s 0 -c 5 -o gpurun_out/r3x_prof -f $C > gpurun_out/r3x_ncu.log 2>&1
tail -n 2 gpurun_out/r3x_ncu.log; tail -3 gpurun_out/r3x_tests.log
