# round 2, call 3B (8 GPUs): the bench as the driver launches it at N = 8 and N = 4
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r3j_n$N.json 2> gpurun_out/r3j_n$N.err; echo "n$N rc $?"
done
python - <<'PY'
import json
for N in (8, 4):
    d=json.loads(open(f"gpurun_out/r3j_n{N}.json").read().strip().splitlines()[-1]); print(N, "%.4g"%d["value"], round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
