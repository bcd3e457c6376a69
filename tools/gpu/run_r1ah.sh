mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -x) 2>&1 | tail -5
for wl in cfg3 cfg3dense cfg5; do
python bench.py --workload $wl --steps 3 --warmup 3 --cpu-structures 10 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in b['stage_ms'].items()}, b['hits_per_step'], b['gpu_launches'], b.get('parity_on_sample'))"
done
