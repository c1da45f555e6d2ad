for ch in 1 2 4; do
for w in c5 c2; do
ATMRT_CHUNKS=$ch python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t1_$w.json 2> gpurun_out/t1_$w.err; tail -3 gpurun_out/t1_$w.err
python -c "
import json; d=json.load(open('gpurun_out/t1_$w.json')); print('chunks $ch $w', d['ms_per_step'], d['stage_ms'])"
done
done
