set -x
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/mg2.json 2> gpurun_out/mg2.err
tail -c 600 gpurun_out/mg2.err
python bench.py --steps 3 --warmup 3 > gpurun_out/mg1.json 2> gpurun_out/mg1.err
tail -c 300 gpurun_out/mg1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/mgref.json 2> gpurun_out/mgref.err
tail -c 300 gpurun_out/mgref.err
