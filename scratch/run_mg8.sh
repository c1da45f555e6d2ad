N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mg_$N.json 2> gpurun_out/mg_$N.err
tail -c 300 gpurun_out/mg_$N.err
python -c "
import json; d=json.load(open('gpurun_out/mg_$N.json')); print($N, d['ms_per_step'], d['value'], d['stage_ms'], d['e2e']['ms_per_step'] if d['e2e'] else None)"
