"""Latency of the reference-arithmetic (FAITHFUL, bit-identical) path on ONE bar, the shape the reference itself runs:
static solve(), time-dependent solve(dt) and XSolver::new (assembly + factorisation) at a few sizes.  Device-resident,
CUDA events.  Next to every GPU figure the CPU port (oracle/, the reference's algorithm on ONE host thread, as the reference
runs it) on the same box: banded Thomas / banded step, and assembly with the quadrature loop.
Usage: python tools/time_small_bar.py [n ...]   (tuning aid; not part of the test suite)"""
import os, sys, json, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as o
import dzahui_b200 as dz
from dzahui_b200 import solvers, params


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # microseconds


def cpu_timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6  # microseconds


def main():
    sizes = [int(float(a)) for a in sys.argv[1:]] or [100, 1000, 10000, 65536]
    out = {}
    opt = solvers.make_options(assembly_mode=solvers.ASSEMBLY_FAITHFUL, solve_mode=solvers.SOLVE_FAITHFUL)
    for n in sizes:
        x = np.linspace(0.0, 1.0, n)
        mesh = dz.Mesh.from_nodes(x)
        reps = max(3, min(200, 2_000_000 // n))
        p = params.DiffussionParams.time_independent().mu(1.0).b(0.5).boundary_conditions(float(os.environ.get('BC0', '1.0')), 15.0).build()
        s = solvers.DiffussionSolverTimeIndependent(p, mesh, 150, options=opt)
        q = params.DiffussionParams.time_dependent().mu(1.0).b(0.5).boundary_conditions(0.0, 1.0).initial_conditions(np.sin(x[1:-1])).build()
        td = solvers.DiffussionSolverTimeDependent(q, mesh, 150, options=opt)
        out[n] = {"static_solve_us": timed(lambda: s.solve_device(0.0), reps), "td_step_us": timed(lambda: td.solve_device(1e-9), reps),
                  "static_rebuild_us": timed(lambda: s.rebuild(), max(3, reps // 4)), "td_rebuild_us": timed(lambda: td.rebuild(), max(3, reps // 4))}
        # the CPU port on one thread (banded storage: the reference's dense Array2 would be far slower beyond a few thousand nodes)
        creps = max(3, min(200, 400_000 // n))
        bands = o.ti_assemble(x, 1.0, 0.5, (1.0, 15.0), 150)
        tdb = o.td_assemble(x, 1.0, 0.5, 150)
        sess = o.EnsembleSession([b[None, :] for b in tdb], o.td_initial_state(np.sin(x[1:-1]), n, (0.0, 1.0))[None, :], (0.0, 1.0))
        k = 20
        out[n].update({"cpu_static_solve_us": cpu_timed(lambda: o.thomas(*bands), creps),
                       "cpu_td_step_us": cpu_timed(lambda: sess.advance(1e-9, k, 1), max(3, creps // 4)) / k,
                       "cpu_static_rebuild_us": cpu_timed(lambda: o.ti_assemble(x, 1.0, 0.5, (1.0, 15.0), 150), max(3, creps // 8)),
                       "cpu_td_rebuild_us": cpu_timed(lambda: o.td_assemble(x, 1.0, 0.5, 150), max(3, creps // 8))})
        sess.close()
        if n <= 4096:  # the reference's own shape: dense n x n Array2 (zero-filled every time) and quad_pair + cos per (row, node)
            lo, di, up, rhs = bands
            dense = np.diag(di) + np.diag(lo[1:], -1) + np.diag(up[:-1], 1)
            out[n].update({"cpu_dense_static_solve_us": cpu_timed(lambda: o.thomas_dense(dense, rhs), max(3, creps // 4)),
                           "cpu_dense_static_rebuild_us": cpu_timed(lambda: o.ti_assemble(x, 1.0, 0.5, (1.0, 15.0), 150, mode=2), 3)})
    try:
        with open("/proc/cpuinfo") as f:
            out["cpu_model"] = next(l.split(":", 1)[1].strip() for l in f if l.startswith("model name"))
    except Exception:
        pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
