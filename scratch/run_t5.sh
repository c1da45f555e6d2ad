python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t5.json 2> gpurun_out/t5.err; tail -3 gpurun_out/t5.err
python -c "
import json; d=json.load(open('gpurun_out/t5.json')); print(d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['ms_per_step'], 'meta', d['e2e_with_meta']['ms_per_step'])"
