# round 2, call 3R: scene tables uploaded only when they change -- the whole GPU suite, one rank of 8 / 4, the full frame, c4, c2
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r3r_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --workload c5 --emulate-ranks 8 > gpurun_out/r3r_e8.json 2>> gpurun_out/r3r_var.err
$B --workload c5 --emulate-ranks 4 > gpurun_out/r3r_e4.json 2>> gpurun_out/r3r_var.err
$B --workload c5 > gpurun_out/r3r_c5.json 2>> gpurun_out/r3r_var.err
$B --workload c4 > gpurun_out/r3r_c4.json 2>> gpurun_out/r3r_var.err
$B --workload c2 > gpurun_out/r3r_c2.json 2>> gpurun_out/r3r_var.err
tail -3 gpurun_out/r3r_tests.log
python - <<'PY'
import json
for f in ("r3r_e8","r3r_e4","r3r_c5","r3r_c4","r3r_c2"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), "%.4g"%d["value"], {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
