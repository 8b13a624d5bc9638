python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2p_tests.txt
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python tools/d2h_probe.py > gpurun_out/r2p_d2h.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2p_launches_create.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > /dev/null 2>&1
