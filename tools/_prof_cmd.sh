python -m pytest tests/test_gpu_onelaunch.py tests/test_gpu_split_bar.py -x -q 2>&1 | tail -5 > gpurun_out/r3g_tests.txt
for r in 1 2 3; do python tools/time_long_solve.py 10000000 4000000 2000000 1000000 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:round(v['ms'],4) for k,v in d.items()})" >> gpurun_out/r3g_time.txt; done
DZB_OL_TRACE=1 DZB_OL_TRACE_FILE=gpurun_out/r3g_ctas.txt python tools/time_long_solve.py 10000000 > /dev/null 2>&1
