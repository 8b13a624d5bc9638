timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r5a_tests.txt
ncu --set full --import-source on --clock-control none -k regex:k_tri_onelaunch -s 20 -c 3 -o gpurun_out/r5a_onelaunch -f python tools/time_long_solve.py 10000000 > gpurun_out/r5a_ncu.log 2>&1
DZB_OL_TRACE=1 DZB_OL_TRACE_FILE=gpurun_out/r5a_ctas.txt timeout 100 python tools/time_long_solve.py 10000000 > /dev/null 2>&1
