mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -20
for g in 1 2 4; do JB_GROUP=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_g$g.json 2> gpurun_out/r1h_bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1h_bench_g$g.json"))
print("G=$g", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"])
PY
done
