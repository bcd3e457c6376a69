mkdir -p gpurun_out
for L in 1 2 3; do
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_expand_lv<$L," -s 3 -c 1 -o gpurun_out/r1s_lv$L python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1s_ncu$L.log 2>&1; tail -2 gpurun_out/r1s_ncu$L.log | cut -c1-200
done
