# round 2, call 4G: the hit plane carries the hit step next to the list slot (shading: one dependent load less): GPU tests, smoke, the ncu
# capture of the c5 stage kernels processed ON the box into profiles/ncu_summary.json, the default bench line (roofline keyed to
# the capture just taken), then the c4 crossing-march capture and the c5 launch list. Summaries come back in gpurun_out/r4a_profiles.
T=r4g
mkdir -p gpurun_out/${T}_profiles
python profiles/source_sha.py > gpurun_out/${T}_sha.txt
SHA=$(cat gpurun_out/${T}_sha.txt)
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc $?"
C5="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C5 > gpurun_out/${T}_plain5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ray_paths_macro|k_terrain_profile|k_sweep_bits|k_hit_normals|k_shade_tiles' -s 0 -c 5 -o gpurun_out/${T}_c5 -f $C5 > gpurun_out/${T}_ncu5.log 2>&1
echo "ncu c5 rc $?"
python profiles/ncu_to_json.py gpurun_out/${T}_c5.ncu-rep c5 $SHA "round 2, bench.py --workload c5 --steps 2 --warmup 1" > gpurun_out/${T}_json5.log 2>&1
ncu -i gpurun_out/${T}_c5.ncu-rep --page raw --csv | python profiles/summarize.py > gpurun_out/${T}_profiles/c5_stage_kernels.body
cp profiles/ncu_summary.json gpurun_out/${T}_profiles/ncu_summary.json
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4g_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", d["e2e"]["ms_per_step"], {k:d["roofline"].get(k) for k in ("kernel","achieved","peak","frac","traffic")})
PY
$C4 > gpurun_out/${T}_plain4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cross_march|k_thresholds' -s 0 -c 2 -o gpurun_out/${T}_c4 -f $C4 > gpurun_out/${T}_ncu4.log 2>&1
echo "ncu c4 rc $?"
python profiles/ncu_to_json.py gpurun_out/${T}_c4.ncu-rep c4 $SHA "round 2, bench.py --workload c4 --steps 2 --warmup 1" > gpurun_out/${T}_json4.log 2>&1
ncu -i gpurun_out/${T}_c4.ncu-rep --page raw --csv | python profiles/summarize.py > gpurun_out/${T}_profiles/c4_cross_march.body
cp profiles/ncu_summary.json gpurun_out/${T}_profiles/ncu_summary.json
rm -f gpurun_out/${T}_c4.ncu-rep   # 64 MiB come back: the c5 report does, the c4 report is summarised above
$C5 > gpurun_out/${T}_plain5b.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_profiles/r02_c5_launches.csv $C5 > gpurun_out/${T}_ncul.log 2>&1
echo "ncu launch list rc $?"
cat gpurun_out/${T}_json5.log gpurun_out/${T}_json4.log | tail -8
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
$B --workload c5 --emulate-ranks 8 > gpurun_out/${T}_e8.json 2>> gpurun_out/${T}_var.err
$B --workload c4 > gpurun_out/${T}_c4.json 2>> gpurun_out/${T}_var.err
$B --workload c2 > gpurun_out/${T}_c2.json 2>> gpurun_out/${T}_var.err
$B --workload c1 > gpurun_out/${T}_c1.json 2>> gpurun_out/${T}_var.err
python - <<'PY'
import json
for w in ("e8","c4","c2","c1"):
    try:
        d=json.loads(open(f"gpurun_out/r4g_{w}.json").read().strip().splitlines()[-1])
        print(w, round(d["ms_per_step"],3), d.get("stage_ms"), {k:round(x,3) for k,x in (d.get("kernel_ms") or {}).items()})
    except Exception as e: print(w, "ERR", e)
PY
