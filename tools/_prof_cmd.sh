for i in 1 2 3; do
python tools/time_long_solve.py 10000000 2>&1 | tail -1 | cut -c1-60
DZB_LIB=$PWD/dzahui_b200/libdzb_prev.so python tools/time_long_solve.py 10000000 2>&1 | tail -1 | cut -c1-60
done
