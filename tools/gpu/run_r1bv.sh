mkdir -p gpurun_out
python tools/gpu/ingest_scaling.py > gpurun_out/r1bv_ingest.json 2> gpurun_out/r1bv_ingest.err; cat gpurun_out/r1bv_ingest.json; tail -2 gpurun_out/r1bv_ingest.err
(time timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_api_gpu.py -m gpu -q) 2>&1 | tail -4
