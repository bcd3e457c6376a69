mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_sharding.py -x -q) 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1aa_n2.json 2> gpurun_out/r1aa_n2.err; tail -1 gpurun_out/r1aa_n2.json | cut -c1-700; tail -3 gpurun_out/r1aa_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
