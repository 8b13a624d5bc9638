ncu --set full --import-source on --clock-control none -k regex:k_tri_onelaunch -s 20 -c 2 -o gpurun_out/r5f_onelaunch -f python tools/time_long_solve.py 10000000 > gpurun_out/r5f_ncu.log 2>&1
