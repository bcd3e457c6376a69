mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_order.py tests/test_api_gpu.py tests/test_cli_gpu.py -m gpu -q -x) 2>&1 | tail -5
python tools/gpu/api_timing.py 2>&1 | tail -7
