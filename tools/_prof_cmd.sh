N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577"
timeout 400 $T tools/split_bar_check.py --nodes 10000000 > gpurun_out/r2u_split${N}.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT DZB_COMM_NO_PEER=1 timeout 400 $T tools/split_bar_check.py --nodes 10000000 --no-check > gpurun_out/r2u_split${N}_nccl.txt 2>&1
timeout 300 $T tools/d2h_probe.py > gpurun_out/r2u_d2h${N}.txt 2>&1
timeout 600 $T bench.py --gpus $N --steps 100 > gpurun_out/r2u_bench${N}.json 2> gpurun_out/r2u_bench${N}.err
