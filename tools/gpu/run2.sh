mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r1d_tests.log 2>&1; tail -5 gpurun_out/r1d_tests.log
for g in 8 4; do JB_GROUP=$g JB_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_bench_g$g.json 2> gpurun_out/r1d_bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1d_bench_g$g.json"))
print("G=$g", d["value"], d["ms_per_step"], "e2e", d["e2e"], d["stage_ms"], d["work"])
PY
grep -A10 "chunk" gpurun_out/r1d_bench_g$g.err | tail -11; done
python tools/gpu/e2e_phases.py cfg3 2>&1 | tail -6
JB_GROUP=4 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 92 -c 3 -o gpurun_out/r1d_expand python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu.log 2>&1; tail -2 gpurun_out/r1d_ncu.log; ls -la gpurun_out/
