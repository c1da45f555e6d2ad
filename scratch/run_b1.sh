set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c5.json 2> gpurun_out/b1_c5.err; tail -c 1500 gpurun_out/b1_c5.json
python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b1_c2.json 2> gpurun_out/b1_c2.err; tail -c 900 gpurun_out/b1_c2.json
nvcc -O3 -arch=sm_100a -o scratch/lat scratch/lat.cu && ./scratch/lat > gpurun_out/lat.txt; cat gpurun_out/lat.txt
