mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', b['ms_per_step'], b['e2e']['ms_per_step'], b['stage_ms'])"; }
run chunk16
JB_LIB_PATH=ab/libjess_chunk32.so run chunk32
JB_GROUP=1 run chunk16_G1
JB_GROUP=4 run chunk16_G4
