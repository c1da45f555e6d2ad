# round 2, call 4E: k_cross_march at 5 (96 registers) and 4 (128 registers) blocks per SM against 6 (80 registers, 1.1 KB of spill loads)
cp atm_raytracer_b200/libatmrt_cuda.so /tmp/base_lib.so
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in base expC expD; do
  if [ $v = base ]; then cp /tmp/base_lib.so atm_raytracer_b200/libatmrt_cuda.so; else cp scratch/${v}_lib.so atm_raytracer_b200/libatmrt_cuda.so; fi
  $B --workload c4 > gpurun_out/r4e_${v}_c4.json 2> gpurun_out/r4e_${v}.err
done
cp /tmp/base_lib.so atm_raytracer_b200/libatmrt_cuda.so
python - <<'PY'
import json
for v in ("base","expC","expD"):
    try:
        d=json.loads(open(f"gpurun_out/r4e_{v}_c4.json").read().strip().splitlines()[-1])
        print(v, round(d["ms_per_step"],3), d.get("stage_ms",{}).get("march"), {k:round(x,3) for k,x in (d.get("kernel_ms") or {}).items()})
    except Exception as e: print(v, "ERR", e)
PY
