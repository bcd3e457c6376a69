mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
JB_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1j_bench.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"], d["work"])
PY
grep -A10 "chunk" gpurun_out/r1j_bench.err | tail -11
python bench.py --steps 10 --warmup 3 | tail -c 1500
