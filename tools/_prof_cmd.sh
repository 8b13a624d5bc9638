set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2m_tests.txt
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python bench.py --impl reference --steps 20 > gpurun_out/r2m_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2m_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_td_step_lean -c 2 -o gpurun_out/r2m_tdstep python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/r2m_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tri_onelaunch -c 2 -o gpurun_out/r2m_onelaunch python tools/time_long_solve.py 10000000 > gpurun_out/r2m_ncu3.log 2>&1
