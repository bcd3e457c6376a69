mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r1ay_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_order.py -m gpu -q -x -k "cfg1 or lnb_t01_2_5_5 or single_mixed or amy_doctest or short_templates or ordering_path or pipelined or mc2_best0_re1 or (record_order and cfg3dense and reference-2)" 2>&1 | tail -4
echo "exit $?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r1ay_memcheck.log; tail -3 gpurun_out/r1ay_memcheck.log
