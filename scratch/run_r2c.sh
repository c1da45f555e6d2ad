# round 2, call C: stage C as bit-mask sweep + hit normals + row shading; stage A walk anchors -- tests, variants, ncu
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
ATMRT_STAGE_C=legacy ATMRT_WALK_ANCHORS=0 $B > gpurun_out/r2c_legacy_noanchor.json 2> gpurun_out/r2c_var.err
ATMRT_WALK_ANCHORS=0 $B > gpurun_out/r2c_new_noanchor.json 2>> gpurun_out/r2c_var.err
for mb in 10 12 16; do ATMRT_SWEEP_MB=$mb $B > gpurun_out/r2c_new_mb$mb.json 2>> gpurun_out/r2c_var.err; done
python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_c2.json 2>> gpurun_out/r2c_var.err
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sweep_bits|k_terrain_profile|k_hit_normals|k_shade_rows' -s 4 -c 4 -o gpurun_out/r2c_prof -f $C > gpurun_out/r2c_ncu.log 2>&1
tail -n 3 gpurun_out/r2c_ncu.log
