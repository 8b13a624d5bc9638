python bench.py > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
python bench.py --impl reference --steps 20 > gpurun_out/r3i_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3i_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r3i_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tri_onelaunch -c 2 -o gpurun_out/r3i_onelaunch python tools/time_long_solve.py 10000000 > gpurun_out/r3i_ncu3.log 2>&1
python tools/time_long_solve.py 300000 1000000 2000000 4000000 10000000 40000000 > gpurun_out/r3i_sizes.json 2>&1
DZB_OL_TRACE=1 DZB_OL_TRACE_FILE=gpurun_out/r3i_ctas.txt python tools/time_long_solve.py 10000000 > /dev/null 2>&1
