set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python bench.py --impl reference --steps 20 > gpurun_out/r2z_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2z_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_td_step_lean -c 2 -o gpurun_out/r2z_tdstep python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/r2z_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_thomas_factor_tile -c 1 -o gpurun_out/r2z_factor python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/r2z_ncu4.log 2>&1
python tools/time_small_bar.py 11 78 1000 10000 65536 > gpurun_out/r2z_smallbar.json 2> gpurun_out/r2z_smallbar.err
