mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) 2>&1 | tail -6
python tools/gpu/api_timing.py 2>&1 | tail -8
