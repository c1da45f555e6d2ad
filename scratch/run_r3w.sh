# round 2, call 3W: one rank of 8, split and unsplit, twice each on one box; the split tests
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "split_frame_with_crossing or frame_split" 2>&1 | tail -3
B="python bench.py --workload c5 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --emulate-ranks 8"
for i in 1 2; do
$B > gpurun_out/r3w_e8_$i.json 2>> gpurun_out/r3w_var.err
$B --sweep-bands 1 > gpurun_out/r3w_e8_nosplit_$i.json 2>> gpurun_out/r3w_var.err
done
python - <<'PY'
import json
for f in ("r3w_e8_1","r3w_e8_nosplit_1","r3w_e8_2","r3w_e8_nosplit_2"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
