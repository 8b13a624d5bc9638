timeout 500 python tools/c4_intended_check.py 100000 1000000 10000000 > gpurun_out/r4e_intended.txt 2> gpurun_out/r4e_intended.err
