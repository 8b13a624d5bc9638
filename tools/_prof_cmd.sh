python -m pytest tests/test_gpu_onelaunch.py tests/test_gpu_split_bar.py -x -q 2>&1 | tail -3 > gpurun_out/r3q_tests.txt
python tools/time_long_solve.py 1000000 4000000 10000000 19000000 2>&1 | tail -1 > gpurun_out/r3q_sizes.json
