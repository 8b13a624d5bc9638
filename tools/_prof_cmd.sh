DZB_OL_TRACE=1 DZB_OL_TRACE_FILE=gpurun_out/r4b_ctas.txt python tools/time_long_solve.py 10000000 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_tri_onelaunch -s 20 -c 1 -o gpurun_out/r4b_onelaunch -f python tools/time_long_solve.py 10000000 > gpurun_out/r4b_ncu.log 2>&1
