mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) 2>&1 | tail -5
python bench.py > gpurun_out/r1ap_bench.json 2> gpurun_out/r1ap_bench.err; cut -c1-250 gpurun_out/r1ap_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1ap_ref.json 2>/dev/null; cut -c1-250 gpurun_out/r1ap_ref.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1ap_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1ap_launch.log 2>&1
