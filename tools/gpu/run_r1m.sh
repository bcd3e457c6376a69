mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=5) 2>&1 | tail -14
JB_DEBUG=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; tail -c 3000 gpurun_out/r1m_bench.json
grep -A20 "chunk" gpurun_out/r1m_bench.err | tail -21
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1m_ref.json 2>&1; tail -c 900 gpurun_out/r1m_ref.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1m_ncu.log 2>&1
tail -1 gpurun_out/r1m_ncu.log | cut -c1-200
