# round 2, call 4F (2 GPUs): `gen --gpus 2` with the overlays (atmrt_group_context on a group of two), the group tests, bench N = 2
timeout 300 python -m pytest tests/test_host_gen.py tests/test_parallel_gloo.py -q -m gpu 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "group or shard" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4f_n2.json 2> gpurun_out/r4f_n2.err; echo "n2 rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4f_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"],3), d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("kernel_ms"))
PY
