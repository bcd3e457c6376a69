mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],2), round(b['stage_ms']['expand_ms'],3))"; }
run base
JB_LIB_PATH=ab/lib_l1_8.so run l1_8
JB_LIB_PATH=ab/lib_deep8_from6.so run deep8_from6
JB_LIB_PATH=ab/lib_deep32_from2.so run deep32_from2
(time timeout 1200 python -m pytest tests -m gpu -q) 2>&1 | tail -5
python bench.py > gpurun_out/r1bj_bench.json 2> gpurun_out/r1bj_bench.err; cut -c1-250 gpurun_out/r1bj_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1bj_ref.json 2>/dev/null; cut -c1-250 gpurun_out/r1bj_ref.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1bj_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1bj_launch.log 2>&1
