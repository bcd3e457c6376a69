mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r1bl_tests.log 2>&1; tail -4 gpurun_out/r1bl_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PYTHONPATH=. python tools/gpu/sites_timing.py 1.0 > gpurun_out/r1bl_sites.json 2> gpurun_out/r1bl_sites.err; cat gpurun_out/r1bl_sites.json; tail -3 gpurun_out/r1bl_sites.err
PYTHONPATH=. python tools/gpu/api_timing.py > gpurun_out/r1bl_api.log 2>&1; tail -15 gpurun_out/r1bl_api.log
python bench.py --no-cpu-baseline > gpurun_out/r1bl_bench.json 2> gpurun_out/r1bl_bench.err; cut -c1-200 gpurun_out/r1bl_bench.json
