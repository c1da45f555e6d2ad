N=$1
nvidia-smi -L | wc -l
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
grep -h "via\|NVLS\|Channel 00" gpurun_out/scale_$N.json gpurun_out/scale_$N.err | head -8
tail -c 300 gpurun_out/scale_$N.err
