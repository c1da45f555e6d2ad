# round 2, call 3A: stage B with 32 steps per macro step at 25 m (800 m, the span 16 x 50 m already uses): tests, c5 bench, emulated rank of 8
python profiles/source_sha.py > gpurun_out/r3a_sha.txt
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r3a_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --workload c5 > gpurun_out/r3a_c5.json 2>> gpurun_out/r3a_var.err
$B --workload c5 --emulate-ranks 8 > gpurun_out/r3a_e8.json 2>> gpurun_out/r3a_var.err
tail -3 gpurun_out/r3a_tests.log
python - <<'PY'
import json
for f in ("r3a_c5.json","r3a_e8.json"):
    d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
