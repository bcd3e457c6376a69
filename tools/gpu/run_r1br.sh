mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decay or trim") > gpurun_out/r1br_decay.log 2>&1; tail -30 gpurun_out/r1br_decay.log | cut -c1-200
for w in cfg4shard cfg5; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r1br_$w.err | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), b['work']['retries'])"; done
