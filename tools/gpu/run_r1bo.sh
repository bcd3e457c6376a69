mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_expand_lv -s 24 -c 3 -o gpurun_out/r1bo_lv123 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1bo_ncu.log 2>&1; tail -1 gpurun_out/r1bo_ncu.log | cut -c1-200
ls -la gpurun_out/
