mkdir -p gpurun_out; rm -f gpurun_out/r1bp_mem.log
(time JB_MEMLOG=gpurun_out/r1bp_mem.log timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r1bp_tests.log 2>&1; tail -4 gpurun_out/r1bp_tests.log
sort -n gpurun_out/r1bp_mem.log | tail -1; tail -3 gpurun_out/r1bp_mem.log
python bench.py > gpurun_out/r1bp_bench.json 2> gpurun_out/r1bp_bench.err; cut -c1-200 gpurun_out/r1bp_bench.json
