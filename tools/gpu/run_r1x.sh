mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_order.py::test_max_candidates_matches_reference) 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1x_bench.json 2>&1; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r1x_bench.json').read().strip().splitlines()[-1])
print('chunk8', b['ms_per_step'], b['e2e']['ms_per_step'], b['stage_ms'], b['gpu_launches'])
PY
JB_LIB_PATH=ab/libjess_chunk16.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk16', b['ms_per_step'], b['e2e']['ms_per_step'], b['stage_ms'])"
