python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?"; wc -l gpurun_out/final_bench.json
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc $?"; wc -l gpurun_out/final_ref.json
