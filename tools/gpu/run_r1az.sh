mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r1az_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_api_gpu.py -m gpu -q -x 2>&1 | tail -2
tail -2 gpurun_out/r1az_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r1az_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg1 or lnb_t01_2_5_5 or short_templates or (synthetic and cfg3-0.04-kw0-2)" 2>&1 | tail -2
tail -3 gpurun_out/r1az_racecheck.log; grep -c "hazard" gpurun_out/r1az_racecheck.log
