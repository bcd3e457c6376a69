(timeout 900 python -m pytest tests/test_gpu_order.py tests/test_api_gpu.py tests/test_cli_gpu.py -m gpu -q -x) 2>&1 | tail -3
python tools/gpu/latency_probe.py
