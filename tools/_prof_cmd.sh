# Scratch command file for `gpurun -- 'bash tools/_prof_cmd.sh'` (rewritten per experiment).  The round-2 records under profiles/ were
# made with these four steps (N = 1), each only after the previous one had exited 0:
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --set full --import-source on --clock-control none -k regex:k_tri_onelaunch -s 20 -c 2 -o gpurun_out/onelaunch -f python tools/time_long_solve.py 10000000 > gpurun_out/ncu.log 2>&1
# then here: tools/ncu_summary.py, tools/ncu_traffic.py, tools/ncu_stalls.py, tools/ncu_lines.py on gpurun_out/onelaunch.ncu-rep
