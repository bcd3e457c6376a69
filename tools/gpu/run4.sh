mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
for g in 2 4; do JB_GROUP=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_g$g.json 2> gpurun_out/r1f_bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1f_bench_g$g.json"))
print("G=$g", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"])
PY
done
JB_DEBUG=1 python tools/gpu/e2e_phases.py cfg3 2>&1 | grep -v "level\|chunk" | tail -12
