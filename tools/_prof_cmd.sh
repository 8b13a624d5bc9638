for s in 0 60 100 140 180; do
echo "skew $s" >> gpurun_out/r3k_skew.txt
DZB_OL_SKEW=$s python tools/time_long_solve.py 10000000 4000000 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:round(v['ms'],4) for k,v in d.items()})" >> gpurun_out/r3k_skew.txt
done
DZB_OL_SKEW=100 python -m pytest tests/test_gpu_onelaunch.py -x -q 2>&1 | tail -2 >> gpurun_out/r3k_skew.txt
