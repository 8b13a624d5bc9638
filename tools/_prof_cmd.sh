python -m pytest tests/test_gpu_split_bar.py -x -q 2>&1 | tail -15 > gpurun_out/r2n_tests.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/r2n_tests.txt
python tools/split_bar_check.py --nodes 1000003 10000000 > gpurun_out/r2n_split1.txt 2>&1
