mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_order.py -m gpu -x -q) 2>&1 | tail -25
(time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_order.py) 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-420
ncu --set full --clock-control none --import-source on -k regex:k_expand_lv -s 24 -c 3 -o gpurun_out/r1t_lv123 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1t_ncu.log 2>&1; tail -1 gpurun_out/r1t_ncu.log | cut -c1-200
