(time python -m pytest tests -m gpu -x -q) > gpurun_out/r1c_tests.log 2>&1; tail -15 gpurun_out/r1c_tests.log
for g in 8 4; do JB_GROUP=$g JB_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_g$g.json 2> gpurun_out/r1c_bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1c_bench_g$g.json"))
print("G=$g", d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"], d["work"])
PY
grep -A19 "chunk" gpurun_out/r1c_bench_g$g.err | tail -20; done
