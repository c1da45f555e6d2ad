python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c5.json 2> gpurun_out/b1_c5.err
python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c2.json 2> gpurun_out/b1_c2.err
CMD="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths' -s 1 -c 1 -o gpurun_out/r1d_paths -f $CMD > gpurun_out/ncu_b.log 2>&1
tail -n 2 gpurun_out/ncu_b.log
