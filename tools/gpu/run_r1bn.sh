mkdir -p gpurun_out; rm -f gpurun_out/r1bn_mem.log
(time JB_MEMLOG=gpurun_out/r1bn_mem.log timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r1bn_tests.log 2>&1; tail -4 gpurun_out/r1bn_tests.log
sort -n gpurun_out/r1bn_mem.log | tail -1; tail -3 gpurun_out/r1bn_mem.log
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r1bn_bench.json 2> gpurun_out/r1bn_bench.err; cut -c1-200 gpurun_out/r1bn_bench.json
