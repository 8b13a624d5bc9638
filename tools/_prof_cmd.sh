set -x
python tools/time_long_solve.py 10000000 > gpurun_out/r2c_time.json 2>gpurun_out/r2c_time.err
ncu --set full --import-source on --clock-control none -k regex:k_tri_onelaunch -c 3 -o gpurun_out/r2c_onelaunch -f python tools/time_long_solve.py 10000000 > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_ncu.log
cat gpurun_out/r2c_time.json
