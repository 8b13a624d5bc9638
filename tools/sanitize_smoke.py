#!/usr/bin/env python3
"""Small pass over every kernel family for `compute-sanitizer --tool memcheck` (ragged sizes included)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dzahui_b200 as dz  # noqa: E402
from dzahui_b200 import distributed as D, synthetic  # noqa: E402

dz.set_device(0)
p = dz.DiffussionParams.time_independent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).build()
for n in (3, 9, 78, 1025, 2049, 4097, 20001):
    mesh = synthetic.regular_bar(n)
    for opt in (dz.make_options(dz.ASSEMBLY_FAITHFUL, dz.SOLVE_FAITHFUL), dz.make_options(dz.ASSEMBLY_FAST, dz.SOLVE_PARALLEL, 1)):
        s = dz.DiffussionSolverTimeIndependent(p, mesh, 150, opt)
        u = s.solve(0.0)
        assert u[0] == 1.0 and u[-1] == 15.0 and np.all(np.isfinite(u)), n
        s.close()
q = dz.StokesParams.static_pressure().hydrostatic_pressure(100.0).density(1.0).force_function(lambda x: -10.0).build()
dz.StokesSolver1D(q, synthetic.regular_bar(11), 350).solve(0.0)
for n in (3, 40, 300, 513, 1024):
    ids = np.arange(5)
    meshes = synthetic.ensemble_meshes(ids, n)
    ic = synthetic.ensemble_initial_conditions(meshes, ids)
    for opt in (dz.make_options(dz.ASSEMBLY_FAST, dz.SOLVE_PARALLEL), dz.make_options(dz.ASSEMBLY_FAST, dz.SOLVE_PARALLEL, -1, True),
                dz.make_options(dz.ASSEMBLY_FAITHFUL, dz.SOLVE_FAITHFUL)):
        e = dz.EnsembleTimeDependent(meshes, 1.0, 1.0, (1.0, 15.0), ic, 150, opt)
        e.step(1e-9, 1)
        e.step(1e-9, 5)
        assert np.all(np.isfinite(e.state()))
        e.close()
ic = np.zeros(20001 - 2)
t = dz.DiffussionParams.time_dependent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).initial_conditions(ic).build()
td = dz.DiffussionSolverTimeDependent(t, synthetic.regular_bar(20001), 150, dz.make_options(dz.ASSEMBLY_FAST, dz.SOLVE_PARALLEL))
td.solve(1e-10)
bar = D.DistributedBarTimeIndependent(p, synthetic.regular_bar(5003), 150, comm=D.LocalComm(), refine_steps=1, blocks_per_process=3)
assert np.all(np.isfinite(bar.solve()))
m = dz.Mesh.builder(os.path.join(ROOT, "tests", "golden", "assets", "1dbar.obj")).build_mesh_1d()
s = dz.DiffussionSolverTimeIndependent(p, m, 150)
m.update_gradient_1d(s.solve_device(0.0))
print("sanitize smoke ok")
