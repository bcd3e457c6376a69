mkdir -p gpurun_out; rm -f gpurun_out/r1bm_mem.log
(time JB_MEMLOG=gpurun_out/r1bm_mem.log timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r1bm_tests.log 2>&1; tail -4 gpurun_out/r1bm_tests.log
awk '{print $1}' gpurun_out/r1bm_mem.log | sort -n | tail -1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
