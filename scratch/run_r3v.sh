# round 2, call 3V: the lower band's counters after the checks -- the whole GPU suite, one rank of 8, the full frame
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r3v_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --emulate-ranks 8 > gpurun_out/r3v_e8.json 2>> gpurun_out/r3v_var.err
$B > gpurun_out/r3v_c5.json 2>> gpurun_out/r3v_var.err
tail -3 gpurun_out/r3v_tests.log
python - <<'PY'
import json
for f in ("r3v_e8","r3v_c5"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), "%.4g"%d["value"], {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
