mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_api_gpu.py -m gpu -q -x) 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:k_expand_lv -s 24 -c 3 -o gpurun_out/r1ae_lv123 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1ae_ncu.log 2>&1; tail -1 gpurun_out/r1ae_ncu.log | cut -c1-200
