# round 2, call 4B: k_shade_tiles at 5 and 6 blocks per SM (48 / 40 registers, small spills) against 4 (64 registers); the GeoTIFF GPU test
cp atm_raytracer_b200/libatmrt_cuda.so /tmp/base_lib.so
timeout 300 python -m pytest tests/test_geotiff.py -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in base exp5 exp6; do
  if [ $v = base ]; then cp /tmp/base_lib.so atm_raytracer_b200/libatmrt_cuda.so; else cp scratch/${v}_lib.so atm_raytracer_b200/libatmrt_cuda.so; fi
  $B --workload c5 > gpurun_out/r4b_${v}_c5.json 2> gpurun_out/r4b_${v}.err
  $B --workload c5 --emulate-ranks 8 > gpurun_out/r4b_${v}_e8.json 2>> gpurun_out/r4b_${v}.err
  $B --workload c2 > gpurun_out/r4b_${v}_c2.json 2>> gpurun_out/r4b_${v}.err
done
cp /tmp/base_lib.so atm_raytracer_b200/libatmrt_cuda.so
python - <<'PY'
import json
for v in ("base","exp5","exp6"):
    for w in ("c5","e8","c2"):
        try:
            d=json.loads(open(f"gpurun_out/r4b_{v}_{w}.json").read().strip().splitlines()[-1])
            print(v, w, round(d["ms_per_step"],3), {k:round(x,3) for k,x in (d.get("kernel_ms") or {}).items()} or d.get("stage_ms"))
        except Exception as e: print(v, w, "ERR", e)
PY
