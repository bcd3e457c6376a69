mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -x) 2>&1 | tail -5
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/r1al_cfg2.json 2>gpurun_out/r1al_cfg2.err; python -c "
import json; b=json.loads(open('gpurun_out/r1al_cfg2.json').read().strip().splitlines()[-1]); print('cfg2', b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], b['hits_per_step'], b['cpu_baseline']['value'], b['parity_on_sample'], b['gpu_launches'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
