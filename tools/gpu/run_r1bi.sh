run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],2), round(b['stage_ms']['expand_ms'],3))"; }
run base
JB_LIB_PATH=ab/lib_l1_8.so run l1_8
JB_LIB_PATH=ab/lib_deep8_from6.so run deep8_from6
JB_LIB_PATH=ab/lib_deep32_from2.so run deep32_from2
