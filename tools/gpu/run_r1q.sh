mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; cut -c1-1500 gpurun_out/r1q_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1q_ref.json 2>&1; cut -c1-600 gpurun_out/r1q_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1q_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand_lv -s 20 -c 4 -o gpurun_out/r1q_expand python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1q_ncu.log 2>&1; tail -1 gpurun_out/r1q_ncu.log | cut -c1-200
