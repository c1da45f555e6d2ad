# round 2, call 3O: sweep ahead of the path check -- the whole GPU suite, one rank of 8, the full frame
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r3o_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --emulate-ranks 8 > gpurun_out/r3o_e8.json 2>> gpurun_out/r3o_var.err
$B --emulate-ranks 4 > gpurun_out/r3o_e4.json 2>> gpurun_out/r3o_var.err
$B > gpurun_out/r3o_c5.json 2>> gpurun_out/r3o_var.err
tail -3 gpurun_out/r3o_tests.log
python - <<'PY'
import json
for f in ("r3o_e8","r3o_e4","r3o_c5"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), "%.4g"%d["value"], {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
