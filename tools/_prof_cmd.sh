timeout 400 python -m pytest tests/test_gpu_onelaunch.py tests/test_gpu_split_bar.py -x -q 2>&1 | tail -3 > gpurun_out/r5d_tests.txt
timeout 200 python tools/time_long_solve.py 1000000 2500000 5000000 10000000 12000007 19000000 2>&1 | tail -1 > gpurun_out/r5d_sizes.json
DZB_OL_PLACEMENT=0 timeout 200 python tools/time_long_solve.py 5000000 10000000 19000000 2>&1 | tail -1 > gpurun_out/r5d_sizes_noplace.json
DZB_OL_TRACE=1 DZB_OL_TRACE_FILE=gpurun_out/r5d_ctas.txt timeout 100 python tools/time_long_solve.py 10000000 > /dev/null 2>&1
