python bench.py --workload c2 --generator Rectilinear --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/rect_c2.json 2> gpurun_out/rect_c2.err; tail -3 gpurun_out/rect_c2.err
python -c "
import json; d=json.load(open('gpurun_out/rect_c2.json')); print('c2', d['ms_per_step'], d['value'], d['ray_steps_per_s'])"
