# round 2, call F: row bands in the sweep, dxr prefetch in stage B -- tests, band variants at N=1 and as one rank of 8
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2f_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for nb in 0 1 2 4; do $B --sweep-bands $nb > gpurun_out/r2f_n1_nb$nb.json 2>> gpurun_out/r2f_var.err; done
for nb in 1 4 8 16 32; do $B --emulate-ranks 8 --sweep-bands $nb > gpurun_out/r2f_e8_nb$nb.json 2>> gpurun_out/r2f_var.err; done
for nb in 1 4 8; do $B --emulate-ranks 2 --sweep-bands $nb > gpurun_out/r2f_e2_nb$nb.json 2>> gpurun_out/r2f_var.err; done
