mkdir -p gpurun_out
for wl in cfg3 cfg3dense cfg5; do
python bench.py --workload $wl --steps 5 --warmup 3 --cpu-structures 10 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl auto', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), b['hits_per_step'], b.get('parity_on_sample'))"
done
for n in 6 8; do for wl in cfg3dense cfg5; do
JB_PIPELINE=$n python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl slices $n e2e', round(b['e2e']['ms_per_step'],2))"
done; done
python bench.py --workload cfg4shard --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4shard auto', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
