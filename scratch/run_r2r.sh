# round 2, call R: crossing march, a block = one band of four adjacent columns
python profiles/source_sha.py > gpurun_out/r2r_sha.txt
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2r_tests.log
timeout 600 python scratch/c4_probe.py > gpurun_out/r2r_c4probe.log 2>&1
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2r_c4.json 2> gpurun_out/r2r_var.err
C="python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > gpurun_out/r2r_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_cross_march|k_thresholds' -s 0 -c 2 -o gpurun_out/r2r_prof -f $C > gpurun_out/r2r_ncu.log 2>&1
tail -n 3 gpurun_out/r2r_ncu.log; cat gpurun_out/r2r_c4probe.log; tail -3 gpurun_out/r2r_tests.log
