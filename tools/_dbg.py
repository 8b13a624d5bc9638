import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dzahui_b200 as gpu
from dzahui_b200 import synthetic
opt = gpu.make_options(gpu.ASSEMBLY_FAST, gpu.SOLVE_PARALLEL, 0)
n = 3_000_000
mesh = synthetic.regular_bar(n)
p = gpu.DiffussionParams.time_independent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).build()
os.environ["DZB_NO_ONELAUNCH"] = "1"
m = gpu.DiffussionSolverTimeIndependent(p, mesh, 150, opt)
um = m.solve(0.0)
del os.environ["DZB_NO_ONELAUNCH"]
s = gpu.DiffussionSolverTimeIndependent(p, mesh, 150, opt)
for k in range(8):
    u = s.solve(0.0)
    e = u - um
    tiles = (n + 2047) // 2048
    print("solve", k, "rel", np.linalg.norm(e) / np.linalg.norm(um), "first bad row", int(np.argmax(np.abs(e) > 1e-9)), "max at", int(np.argmax(np.abs(e))), float(np.abs(e).max()))
    # error at CTA-range interface rows: rows tb0*2048 and tb1*2048-1
    units = tiles + 1; grid = 296
    tb = [min(tiles, b * units // grid) for b in range(grid)] + [tiles]
    ends = [min(n - 1, tb[b] * 2048) for b in range(0, grid, 37)]
    print("   err at some CTA starts:", [(r, float(e[r])) for r in ends])
    print("   err sample every 300k:", [float(e[i]) for i in range(0, n, 300000)])
