mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
for w in "cfg5 0.4" "cfg4shard 0.2" "cfg3dense 1.0"; do set -- $w; JB_DEBUG=1 python bench.py --workload $1 --scale $2 --steps 3 --warmup 3 --cpu-structures 20 > gpurun_out/r1k_bench_$1.json 2> gpurun_out/r1k_bench_$1.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1k_bench_$1.json"))
print("$1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"], d["work"], d.get("cpu_baseline"), d.get("parity_on_sample"))
PY
grep "chunk" gpurun_out/r1k_bench_$1.err | tail -3; grep -A12 "chunk" gpurun_out/r1k_bench_$1.err | tail -12; done
