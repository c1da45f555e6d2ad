# round 2, call 4I: the default bench line of the final tree (bench.py with the config keys of both arms aligned)
python profiles/source_sha.py > gpurun_out/r4i_sha.txt
python bench.py > gpurun_out/r4i_bench.json 2> gpurun_out/r4i_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4i_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", d["e2e"]["ms_per_step"], d["config"].keys(), d["roofline"]["frac"], d["march_mode"][:20])
PY
