mkdir -p gpurun_out
JB_DEBUG=1 python bench.py --workload cfg4shard --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/r1be_debug.err | cut -c1-120
grep -A19 "chunk \[" gpurun_out/r1be_debug.err | tail -20
