mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -8
for cfg in "1 0" "2 0" "2 1"; do set -- $cfg; JB_GROUP=$1 JB_LEGACY=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1o_bench_g$1_leg$2.json 2> gpurun_out/r1o_bench_g$1_leg$2.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1o_bench_g$1_leg$2.json"))
print("G=$1 legacy=$2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"], d["work"]["hits"])
PY
done
JB_DEBUG=1 JB_GROUP=2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep -A12 "chunk" | tail -13
JB_DEBUG=1 JB_GROUP=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep -A12 "chunk" | tail -13
