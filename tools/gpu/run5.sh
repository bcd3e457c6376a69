mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
for g in 1 2; do JB_GROUP=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_g$g.json 2> gpurun_out/r1g_bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/r1g_bench_g$g.json"))
print("G=$g", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"])
PY
done
# DRAM traffic + time of every kernel of one resident step (launch list with dram bytes)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_ncu.log 2>&1
tail -1 gpurun_out/r1g_ncu.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 92 -c 3 -o gpurun_out/r1g_expand python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_ncu2.log 2>&1; tail -1 gpurun_out/r1g_ncu2.log
