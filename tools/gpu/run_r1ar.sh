mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1ar_n8.json 2> gpurun_out/r1ar_n8.err; tail -1 gpurun_out/r1ar_n8.json | cut -c1-900; tail -3 gpurun_out/r1ar_n8.err | cut -c1-300
