# round 2, call 3N: frame split below the horizon -- tests, one rank of 8 and of 4 with and without the split, the full frame
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "frame_split or row_bands or full_size_c5_properties or march_variants or group_" 2>&1 | tail -12 > gpurun_out/r3n_tests.log
B="python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --emulate-ranks 8 > gpurun_out/r3n_e8.json 2>> gpurun_out/r3n_var.err
$B --emulate-ranks 8 --sweep-bands 1 > gpurun_out/r3n_e8_nosplit.json 2>> gpurun_out/r3n_var.err
$B --emulate-ranks 4 > gpurun_out/r3n_e4.json 2>> gpurun_out/r3n_var.err
$B --emulate-ranks 4 --sweep-bands 1 > gpurun_out/r3n_e4_nosplit.json 2>> gpurun_out/r3n_var.err
$B > gpurun_out/r3n_c5.json 2>> gpurun_out/r3n_var.err
tail -4 gpurun_out/r3n_tests.log
python - <<'PY'
import json
for f in ("r3n_e8","r3n_e8_nosplit","r3n_e4","r3n_e4_nosplit","r3n_c5"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
