# round 2, call B: the fused stage C (sweep + shading in one kernel, row bands) -- tests, variants, one ncu capture
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2b_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
ATMRT_STAGE_C=legacy $B > gpurun_out/r2b_legacy.json 2> gpurun_out/r2b_legacy.err
for mb in 6 8 4; do for nb in 1 2 4; do
  ATMRT_FUSED_MB=$mb ATMRT_SWEEP_BANDS=$nb $B > gpurun_out/r2b_mb${mb}_nb${nb}.json 2>> gpurun_out/r2b_var.err
done; done
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sweep_fused|k_terrain_profile|k_rgb_rows' -s 3 -c 3 -o gpurun_out/r2b_prof -f $C > gpurun_out/r2b_ncu.log 2>&1
tail -n 3 gpurun_out/r2b_ncu.log
