python -m pytest tests/test_gpu_onelaunch.py -q 2>&1 | tail -4
python tools/_dbg.py 2>/dev/null | grep solve | cut -c1-50 | sort | uniq -c
python tools/time_long_solve.py 1000000 10000000 2>&1 | tail -1 | cut -c1-500
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
