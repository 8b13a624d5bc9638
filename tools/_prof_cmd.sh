python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2v_tests.txt
python tools/time_small_bar.py 11 78 100 1000 10000 65536 > gpurun_out/r2v_smallbar.json 2> gpurun_out/r2v_smallbar.err
