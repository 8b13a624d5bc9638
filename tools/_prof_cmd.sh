N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577"
timeout 300 python -m pytest tests/test_gpu_split_bar.py -x -q 2>&1 | tail -3 > gpurun_out/r4s_tests${N}.txt
timeout 400 $T tools/split_bar_check.py --nodes 10000000 > gpurun_out/r4s_split${N}.txt 2>&1
timeout 600 $T bench.py --gpus $N --steps 100 > gpurun_out/r4s_bench${N}.json 2> gpurun_out/r4s_bench${N}.err
