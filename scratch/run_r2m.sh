# round 2, call M: Spline temperature functions + regular-terrain path + normals (128,8): tests, default bench, reference arm, ncu capture
python profiles/source_sha.py > gpurun_out/r2m_sha.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2m_tests.log
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo "ref rc $?"
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --emulate-ranks 8 > gpurun_out/r2m_e8.json 2>> gpurun_out/r2m_var.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2m_c4.json 2>> gpurun_out/r2m_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2m_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths_macro|k_terrain_profile|k_sweep_bits|k_hit_normals|k_shade_tiles' -s 0 -c 5 -o gpurun_out/r2m_prof -f $C > gpurun_out/r2m_ncu.log 2>&1
tail -n 3 gpurun_out/r2m_ncu.log
