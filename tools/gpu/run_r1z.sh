mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r1z_bench.json 2> gpurun_out/r1z_bench.err; cut -c1-300 gpurun_out/r1z_bench.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1z_launch.log 2>&1
for wl in cfg3dense cfg5 cfg4shard; do
  (time python bench.py --workload $wl --steps 3 --warmup 3 --cpu-structures 20 > gpurun_out/r1z_$wl.json 2> gpurun_out/r1z_$wl.err) 2>&1 | grep real
  python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r1z_$wl.json').read().strip().splitlines()[-1])
    print('$wl', b['value'], b['ms_per_step'], b['e2e']['value'], b['config'].get('pairs_per_step'), b['hits_per_step'], b.get('cpu_baseline',{}).get('value'), b.get('parity_on_sample'), b['work']['retries'], b['work']['chunks'])
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/r1z_$wl.err').read()[-800:])
PY
done
