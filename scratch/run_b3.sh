python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c5.json 2> gpurun_out/b1_c5.err
python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c2.json 2> gpurun_out/b1_c2.err
bash scratch/ncu_l.sh
