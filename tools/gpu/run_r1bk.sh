mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_api_gpu.py -m gpu -q) > gpurun_out/r1bk_tests.log 2>&1; tail -5 gpurun_out/r1bk_tests.log
PYTHONPATH=. python tools/gpu/sites_timing.py 1.0 > gpurun_out/r1bk_sites.json 2> gpurun_out/r1bk_sites.err; cat gpurun_out/r1bk_sites.json; tail -3 gpurun_out/r1bk_sites.err
