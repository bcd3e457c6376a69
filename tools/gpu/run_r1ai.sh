mkdir -p gpurun_out
(time JB_PIPELINE=3 timeout 1200 python -m pytest tests -m gpu -q -x) 2>&1 | tail -5
(time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_api_gpu.py -m gpu -q -x) 2>&1 | tail -3
for wl in cfg3 cfg3dense cfg5; do
python bench.py --workload $wl --steps 5 --warmup 3 --cpu-structures 10 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), b['hits_per_step'], b.get('parity_on_sample'))"
JB_PIPELINE=1 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl nopipe', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
done
JB_PIPELINE=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 2 slices e2e', round(b['e2e']['ms_per_step'],2))"
JB_PIPELINE=8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 8 slices e2e', round(b['e2e']['ms_per_step'],2))"
