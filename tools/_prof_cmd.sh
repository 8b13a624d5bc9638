N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577"
timeout 200 python -m pytest tests/test_gpu_split_bar.py -x -q 2>&1 | tail -3 > gpurun_out/r4t_tests${N}.txt
timeout 300 $T tools/split_bar_check.py --nodes 10000000 > gpurun_out/r4t_split${N}.txt 2>&1
timeout 400 $T bench.py --gpus $N --steps 100 > gpurun_out/r4t_bench${N}.json 2> gpurun_out/r4t_bench${N}.err
