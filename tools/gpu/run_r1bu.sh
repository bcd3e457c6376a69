mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r1bu_tests.log 2>&1; tail -3 gpurun_out/r1bu_tests.log
python bench.py > gpurun_out/r1bu_bench.json 2> gpurun_out/r1bu_bench.err; cut -c1-220 gpurun_out/r1bu_bench.json
