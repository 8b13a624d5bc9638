"""How far the reference's own f64 result and this library's two assemblies are from the system the reference means to solve.

For the regular bar of BASELINE config 4 (mu = b = 1, bc (1, 15), G = 150) at several sizes:
  truth  = the reference's static diffusion system evaluated and solved in exact (__float128) arithmetic on its f64 inputs
           (oracle: dzo_ti_intended_f128; no band entry is ever rounded to f64),
  ref    = the oracle's f64 assembly + solve_by_thomas (what the reference returns),
  auto   = the library's default mode (bands bit-identical to the reference's, partitioned solve),
  fast   = closed-form rows (one launch at these sizes).
Prints one JSON object per size with the relative L2 distances.  Needs a GPU; tuning / documentation aid.
Usage: python tools/c4_intended_check.py [--jitter] [n ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dzahui_b200 as dz  # noqa: E402
from dzahui_b200 import params, solvers  # noqa: E402
import oracle_lib as o  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    argv = [a for a in sys.argv[1:] if a != "--jitter"]
    jitter = "--jitter" in sys.argv[1:]   # element lengths within 0.4 % of each other (tests/test_gpu_onelaunch.py::jittered) instead of equal
    sizes = [int(float(a)) for a in argv] or [100_000, 1_000_000, 10_000_000]
    threads = max(1, len(os.sched_getaffinity(0)))
    for n in sizes:
        if jitter:
            x = np.concatenate([[0.0], np.cumsum(np.random.default_rng(5).uniform(0.996, 1.004, n - 1))]) / n
        else:
            x = np.linspace(0.0, 1.0, n)
        t0 = time.time()
        truth = o.ti_intended_f128(x, 1.0, 1.0, (1.0, 15.0), 150, threads=threads)
        t_truth = time.time() - t0
        lo, di, up, rhs = o.ti_assemble(x, 1.0, 1.0, (1.0, 15.0), 150, threads=threads)
        ref = o.thomas(lo, di, up, rhs)
        ref_exact_solve = o.thomas(lo, di, up, rhs, kind="f128")
        mesh = dz.Mesh.from_nodes(x)
        p = params.DiffussionParams.time_independent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).build()
        out = {"n": n, "mesh": "jittered 0.4 %" if jitter else "regular", "truth_seconds": round(t_truth, 2),
               "relL2_reference_f64_vs_truth": rel(ref, truth),
               "relL2_reference_bands_solved_in_f128_vs_truth": rel(ref_exact_solve, truth)}
        for name, mode in (("auto", solvers.ASSEMBLY_AUTO), ("fast", solvers.ASSEMBLY_FAST)):
            s = solvers.DiffussionSolverTimeIndependent(p, mesh, 150, options=solvers.make_options(assembly_mode=mode))
            u = np.asarray(s.solve(0.0))
            out[f"relL2_gpu_{name}_vs_truth"] = rel(u, truth)
            out[f"relL2_gpu_{name}_vs_reference_f64"] = rel(u, ref)
            out[f"path_{name}"] = s.path()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
