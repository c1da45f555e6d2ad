python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/mg_2.json 2> gpurun_out/mg_2.err
echo "rc $?"; wc -l gpurun_out/mg_2.json
python -c "
import json; d=json.loads(open('gpurun_out/mg_2.json').read()); print(2, d['ms_per_step'], d['value'], d['stage_ms'], d['e2e']['ms_per_step'] if d['e2e'] else None)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/mg_2ref.json 2> gpurun_out/mg_2ref.err
echo "ref rc $?"; head -c 200 gpurun_out/mg_2ref.json
