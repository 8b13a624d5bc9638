timeout 80 python -m pytest tests/test_gpu_onelaunch.py -x -q -k "c4_parameters or exact_arithmetic" 2>&1 | tail -3 > gpurun_out/r5g_tests.txt
