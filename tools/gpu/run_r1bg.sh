run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],2), round(b['stage_ms']['expand_ms'],3))"; }
run default_1
JB_GROUPS=1,1 run g11
JB_GROUPS=1,1,1 run g111
JB_GROUPS=1,2,2,4,4,4 run g122444
JB_GROUPS=1,2,2,2,2,4,4,4 run g12222444
JB_GROUPS=1,1,2,2,2,2,4,4 run g11222244
