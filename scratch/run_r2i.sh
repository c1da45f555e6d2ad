# round 2, call I: stage B as chain + elements; per-kernel timing; new roofline JSON
python profiles/source_sha.py > gpurun_out/r2i_sha.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r2i_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r2i_n1.json 2>> gpurun_out/r2i_var.err
$B --emulate-ranks 8 > gpurun_out/r2i_e8.json 2>> gpurun_out/r2i_var.err
$B --emulate-ranks 2 > gpurun_out/r2i_e2.json 2>> gpurun_out/r2i_var.err
python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_c2.json 2>> gpurun_out/r2i_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_chain|k_ray_elements|k_terrain_profile|k_sweep_bits|k_hit_normals|k_shade_tiles' -s 0 -c 6 -o gpurun_out/r2i_prof -f $C > gpurun_out/r2i_ncu.log 2>&1
tail -n 3 gpurun_out/r2i_ncu.log
