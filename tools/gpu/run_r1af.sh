mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1af_cfg5_launches.csv python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1af.log 2>&1
python tools/summarise_launches.py gpurun_out/r1af_cfg5_launches.csv | sort -t'|' -k4 -n -r | head -16
