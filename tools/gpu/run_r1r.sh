mkdir -p gpurun_out
JB_DEBUG=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/r1r_debug.err | cut -c1-200
tail -40 gpurun_out/r1r_debug.err
for L in 1 2 3; do
ncu --set full --clock-control none --import-source on -k regex:"k_expand_lv<$L" -s 3 -c 1 -o gpurun_out/r1r_lv$L python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1r_ncu$L.log 2>&1; tail -1 gpurun_out/r1r_ncu$L.log | cut -c1-200
done
