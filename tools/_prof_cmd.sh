timeout 300 python tools/c4_intended_check.py --jitter 100000 1000000 10000000 > gpurun_out/r4f_intended_jitter.txt 2> gpurun_out/r4f_intended.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r4f_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4f_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err
