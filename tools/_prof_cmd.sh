N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577"
timeout 400 $T tools/split_bar_check.py --nodes 10000000 > gpurun_out/r3s_split${N}.txt 2>&1
timeout 600 $T bench.py --gpus $N --steps 100 > gpurun_out/r3s_bench${N}.json 2> gpurun_out/r3s_bench${N}.err
