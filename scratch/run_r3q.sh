# round 2, call 3Q: the default bench on the committed profiles (roofline keyed to the same source sha), smoke()
python profiles/source_sha.py > gpurun_out/r3q_sha.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3q_smoke.log 2>&1; echo "smoke rc $?"
python bench.py > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err; echo "bench rc $?"
python -m pytest tests/test_golden.py -q -m gpu 2>&1 | tail -3
tail -3 gpurun_out/r3q_smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3q_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", d["e2e"]["ms_per_step"], {k:d["roofline"].get(k) for k in ("kernel","achieved","peak","frac","traffic")})
print({k:round(v["frac"],3) if v.get("frac") else None for k,v in d["roofline_kernels"].items()})
PY
