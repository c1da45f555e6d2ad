# round 2, call 3E: atmrt_group_render_tiles -- tests, the default bench (end to end through the one call)
python profiles/source_sha.py > gpurun_out/r3e_sha.txt
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r3e_tests.log
python bench.py > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc $?"
tail -3 gpurun_out/r3e_tests.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3e_bench.json").read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), "e2e", d["e2e"]["ms_per_step"], "with meta", d["e2e_with_meta"]["ms_per_step"], "gen", d["e2e_gen"]["wall_s"])
PY
