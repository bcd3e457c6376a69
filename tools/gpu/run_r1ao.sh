mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_order.py -m gpu -q -x) 2>&1 | tail -5
python tools/gpu/api_timing.py 2>&1 | sed -n 1,3p
JB_KD_UNSORTED=1 python tools/gpu/api_timing.py 2>&1 | sed -n 2,2p
JB_PIPELINE=1 python tools/gpu/api_timing.py 2>&1 | sed -n 1,2p
JB_PIPELINE=1 JB_KD_UNSORTED=1 python tools/gpu/api_timing.py 2>&1 | sed -n 2,2p
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 fused', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), b['stage_ms'], b['gpu_launches'])"
JB_NO_ROOT_FUSION=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 unfused', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
