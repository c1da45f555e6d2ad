# round 2, call 4C: k_shade_tiles with padded staging rows (no 32-way bank conflict) at 4 (A) and 6 (B) blocks per SM; parity of the
# variant against the oracle (golden + c5 properties) for the one that will be kept
cp atm_raytracer_b200/libatmrt_cuda.so /tmp/base_lib.so
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in expA expB; do
  cp scratch/${v}_lib.so atm_raytracer_b200/libatmrt_cuda.so
  $B --workload c5 > gpurun_out/r4c_${v}_c5.json 2> gpurun_out/r4c_${v}.err
  $B --workload c5 --emulate-ranks 8 > gpurun_out/r4c_${v}_e8.json 2>> gpurun_out/r4c_${v}.err
  $B --workload c2 > gpurun_out/r4c_${v}_c2.json 2>> gpurun_out/r4c_${v}.err
  timeout 300 python -m pytest tests/test_golden.py -q -m gpu 2>&1 | tail -2
done
cp /tmp/base_lib.so atm_raytracer_b200/libatmrt_cuda.so
python - <<'PY'
import json
for v in ("expA","expB"):
    for w in ("c5","e8","c2"):
        try:
            d=json.loads(open(f"gpurun_out/r4c_{v}_{w}.json").read().strip().splitlines()[-1])
            print(v, w, round(d["ms_per_step"],3), {k:round(x,3) for k,x in (d.get("kernel_ms") or {}).items()})
        except Exception as e: print(v, w, "ERR", e)
PY
