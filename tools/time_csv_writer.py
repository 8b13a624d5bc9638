"""Times Writer::write's replacement (dzb_csv_write) on 1e7 values.  Usage: python tools/time_csv_writer.py"""
import time, numpy as np, os, tempfile, sys
sys.path.insert(0, os.getcwd())
import dzahui_b200 as dz
d=tempfile.mkdtemp()
w=dz.Writer(d, "snap", ["v_x"], erase_prev_dir=False)
vals=np.random.default_rng(1).uniform(-1,1,10_000_000)
t0=time.perf_counter(); w.write(1.0, vals); dt=time.perf_counter()-t0
f=os.listdir(d); print("csv 1e7 values", os.path.getsize(os.path.join(d,f[0]))/1e6,"MB", round(dt,3),"s")
