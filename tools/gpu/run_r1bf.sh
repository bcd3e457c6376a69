run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],2), round(b['stage_ms']['expand_ms'],3))"; }
run base
JB_GROUP_L1=1 run L1_G1
JB_GROUP_L1=4 run L1_G4
