"""Times one long bar at several sizes (device-resident, CUDA events): with the default build the closed-form assembly is
fused into the partitioned solve (one launch per assemble + solve, onelaunch_kernels.cuh); DZB_NO_ONELAUNCH=1 times the
partitioned solve of materialised bands alone (bigsys_kernels.cuh).
Usage: python tools/time_long_solve.py [n ...]   (tuning aid; not part of the test suite)"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dzahui_b200 as dz
from dzahui_b200 import solvers, params

def main():
    sizes = [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 2_000_000, 10_000_000]
    out = {}
    for n in sizes:
        x = np.linspace(0.0, 1.0, n)
        mesh = dz.Mesh.from_nodes(x)
        p = params.DiffussionParams.time_independent().mu(1.0).b(0.5).boundary_conditions(0.0, 1.0).build()
        s = solvers.DiffussionSolverTimeIndependent(p, mesh, 150, options=solvers.make_options(
            assembly_mode=solvers.ASSEMBLY_FAST, solve_mode=solvers.SOLVE_PARALLEL, refine_steps=0))
        for _ in range(5):
            s.solve_device(0.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            s.solve_device(0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[n] = {"ms": ms, "ns_per_row": ms * 1e6 / n, "GBps_40B": 40.0 * n / ms / 1e6, "GBps_80B": 80.0 * n / ms / 1e6, "path": s.path()}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
