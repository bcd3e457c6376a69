mkdir -p gpurun_out; rm -f gpurun_out/r1bs_mem.log
(time JB_MEMLOG=gpurun_out/r1bs_mem.log timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r1bs_tests.log 2>&1; tail -4 gpurun_out/r1bs_tests.log
sort -n gpurun_out/r1bs_mem.log | tail -1; tail -2 gpurun_out/r1bs_mem.log
python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r1bs_cfg5.err | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2), b['work']['retries'])"
python bench.py > gpurun_out/r1bs_bench.json 2> gpurun_out/r1bs_bench.err; cut -c1-200 gpurun_out/r1bs_bench.json
