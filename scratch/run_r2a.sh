# round 2, call A: GPU tests on the inherited kernels + the new full-size parity tests, smoke, and a c5 bench for reference
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python -m pytest tests -q -m gpu -k "full_size or pixel_angles" -s 2>&1 | grep PARITY > gpurun_out/r2a_parity.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
nproc > gpurun_out/r2a_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r2a_host.txt
