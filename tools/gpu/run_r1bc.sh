(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x) 2>&1 | grep -E "^E|passed|failed|Error" | head -20
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
