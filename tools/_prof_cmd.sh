T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29577"
timeout 600 python -m pytest tests/test_gpu_split_bar.py -x -q 2>&1 | tail -8 > gpurun_out/r2x_tests.txt
timeout 600 $T bench.py --gpus 2 --workload c4 --assembly fast --steps 50 > gpurun_out/r2x_bench_c4_n2.json 2> gpurun_out/r2x_bench_c4_n2.err
timeout 600 python bench.py --workload c4 --assembly fast --steps 50 > gpurun_out/r2x_bench_c4_n1.json 2> gpurun_out/r2x_bench_c4_n1.err
