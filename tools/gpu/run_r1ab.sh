mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q) 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('owner', b['ms_per_step'], b['e2e']['ms_per_step'], b['stage_ms'])"
