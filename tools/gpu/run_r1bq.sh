mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decay or trim") 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_superpose -s 4 -c 1 -o gpurun_out/r1bq_superpose python bench.py --workload cfg3dense --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1bq_ncu.log 2>&1; tail -1 gpurun_out/r1bq_ncu.log | cut -c1-300
