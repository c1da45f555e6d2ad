# round 2, call H: group API (gen --gpus, bench e2e), full default bench
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2h_tests.log
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"
tail -3 gpurun_out/r2h_bench.err
