# round 2, call K: stage A in column chunks with the sweep pipelined behind them
python -m pytest tests -x -q -m gpu -k "bands or full_size_c5 or march_variants or group" 2>&1 | tail -8 > gpurun_out/r2k_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for nc in 1 2 4 8; do $B --column-chunks $nc > gpurun_out/r2k_n1_nc$nc.json 2>> gpurun_out/r2k_var.err; done
for nc in 1 2 4; do $B --emulate-ranks 8 --column-chunks $nc > gpurun_out/r2k_e8_nc$nc.json 2>> gpurun_out/r2k_var.err; done
for nc in 1 4; do python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --column-chunks $nc > gpurun_out/r2k_c2_nc$nc.json 2>> gpurun_out/r2k_var.err; done
