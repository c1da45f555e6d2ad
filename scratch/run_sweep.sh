python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "1 1" "8 8" "8 1" "8 0" "8 2" "16 1" "16 2" "4 1" "32 2"; do
  set -- $cfg
  ATMRT_COLUMN_CHUNKS=$1 ATMRT_EARLY_CHUNKS=$2 python bench.py --workload c5 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/sw.json"))
    print("chunks",sys.argv[1],"early",sys.argv[2],"ms/step",round(d["ms_per_step"],2),{k:round(v,1) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("ERR",e,open("gpurun_out/sw.err").read()[-500:])
PY
done
