mkdir -p gpurun_out
python tools/gpu/ingest_scaling.py > gpurun_out/r1bv_ingest.json 2> gpurun_out/r1bv_ingest.err; cat gpurun_out/r1bv_ingest.json; tail -2 gpurun_out/r1bv_ingest.err
(time timeout 600 python -m pytest tests -m gpu -q) > gpurun_out/r1bv_tests.log 2>&1; tail -4 gpurun_out/r1bv_tests.log
