timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5b_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/r5b_bench.json 2> gpurun_out/r5b_bench.err
timeout 300 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/r5b_ref.json 2> gpurun_out/r5b_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r5b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r5b_ncu_bench.log 2>&1
