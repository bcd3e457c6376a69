mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r1ba_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg1 or lnb_t01_2_5_5 or short_templates or single_mixed or (synthetic and cfg3-0.04-kw0-2)" 2>&1 | tail -2
tail -2 gpurun_out/r1ba_racecheck.log; grep -c "hazard" gpurun_out/r1ba_racecheck.log
python tools/gpu/latency_probe.py
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['ms_per_step'],2))"
