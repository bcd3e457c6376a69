mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_order.py::test_max_candidates_matches_reference) 2>&1 | tail -8
JB_DEBUG=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/r1u_debug.err | cut -c1-300
grep -A12 "chunk \[" gpurun_out/r1u_debug.err | tail -13
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1u_bench.json 2>&1; cut -c1-900 gpurun_out/r1u_bench.json
