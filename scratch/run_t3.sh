python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for w in c5 c2 c4; do
python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t1_$w.json 2> gpurun_out/t1_$w.err; tail -3 gpurun_out/t1_$w.err
python -c "
import json; d=json.load(open('gpurun_out/t1_$w.json')); print('$w', d['ms_per_step'], d['stage_ms'])"
done
