for w in c2 c1; do
python bench.py --workload $w --generator Rectilinear --steps 2 --warmup 1 --no-e2e > gpurun_out/rect_$w.json 2> gpurun_out/rect_$w.err; tail -3 gpurun_out/rect_$w.err
python -c "
import json; d=json.load(open('gpurun_out/rect_$w.json')); print('$w', d['ms_per_step'], d['value'], d['ray_steps_per_s'], d['stage_ms'], d['cpu_baseline'])"
done
CMD="python bench.py --workload c2 --generator Rectilinear --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_rectilinear' -s 1 -c 1 -o gpurun_out/r1_rect -f $CMD > gpurun_out/ncu_rect.log 2>&1
tail -2 gpurun_out/ncu_rect.log
