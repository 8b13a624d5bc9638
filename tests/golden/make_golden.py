#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json and tests/golden/assets/*.obj.

Run in the build container:  python tests/golden/make_golden.py [/root/reference]
 * copies the .obj mesh fixtures that BASELINE.json's configs name out of the reference's assets/ (data, not source);
 * evaluates the CPU oracle (oracle/, a line-faithful restatement of the reference — the Rust crate itself cannot be
   compiled in this image) on configs C1-C3, on the reference's unit-test meshes and on small seeded cases, and stores
   inputs + outputs.  Python's json round-trips f64 exactly (repr).
The golden values double as a regression pin of the oracle itself (tests/test_oracle.py)."""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as o  # noqa: E402
from dzahui_b200 import synthetic  # noqa: E402

ASSETS = ["1dbar.obj", "1dbar_irregular.obj", "1dbar_irregular_small.obj", "1dbar_many_divisions.obj",
          "1dbar_many_divisions_irregular.obj", "big_mesh.obj", "test.obj", "trapezoid.obj", "untitled.obj"]


def L(a):
    return [float(v) for v in np.asarray(a).ravel()]


def big_mesh_nodes(path):
    """C3 projection rule: x token of every `v` line, de-duplicated, ascending."""
    xs = set()
    for line in open(path):
        if line.startswith("v "):
            xs.add(float(line.split(" ")[1]))
    return np.array(sorted(xs))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    adir = os.path.join(HERE, "assets")
    os.makedirs(adir, exist_ok=True)
    if os.path.isdir(os.path.join(ref, "assets")):
        for a in ASSETS:
            shutil.copyfile(os.path.join(ref, "assets", a), os.path.join(adir, a))
    g = {}

    # ---- ingestion
    ing = {}
    for a, mult in [("1dbar.obj", None), ("1dbar_irregular.obj", 3.0), ("1dbar_irregular_small.obj", None),
                    ("1dbar_many_divisions.obj", None), ("1dbar_many_divisions_irregular.obj", 50.0)]:
        m = o.mesh_ingest_1d(os.path.join(adir, a), mult)
        ing[a] = dict(mult=mult, n=m["n"], vertices=L(m["vertices"]), indices=[int(v) for v in m["indices"]],
                      max_length=m["max_length"], prom_width=m["prom_width"], translation=L(m["translation"]),
                      nodes=L(o.filter_for_solving_1d(m["vertices"])))
    g["ingest"] = ing

    # ---- 2D ingestion: build_mesh_2d on the planar meshes the reference ships
    ing2 = {}
    for a in ("big_mesh.obj", "test.obj", "trapezoid.obj", "untitled.obj"):
        try:
            m = o.mesh_ingest_2d(os.path.join(adir, a))
        except o.OracleError as e:
            ing2[a] = dict(error=e.code)
            continue
        ing2[a] = dict(n_vertices=m["n_vertices"], n_triangles=m["n_triangles"], vertices=L(m["vertices"]),
                       indices=[int(v) for v in m["indices"]], boundary_indices=[int(v) for v in m["boundary_indices"]],
                       max_length=m["max_length"], translation=L(m["translation"]))
    g["ingest2d"] = ing2

    # ---- quadrature
    g["quad"] = {str(G): dict(x=L(o.quad_table(G)[0]), w=L(o.quad_table(G)[1])) for G in (2, 3, 7, 50, 99, 100, 101, 150, 350, 600)}

    # ---- C1 hydrostatic pressure (src/bin/static_pressure.rs:6-11)
    mesh = np.array(ing["1dbar.obj"]["nodes"])
    lo, di, up, rhs = o.stokes_assemble(mesh, 1.0, 100.0, 350, -10.0)
    g["C1"] = dict(mesh=L(mesh), rho=1.0, pressure=100.0, force=-10.0, G=350, lo=L(lo), di=L(di), up=L(up), rhs=L(rhs),
                   solution=L(o.thomas(lo, di, up, rhs)))
    # a non-constant force through the same path
    f = lambda x: 3.0 * x * x - 2.0 * x + 0.5
    lo, di, up, rhs = o.stokes_assemble(mesh, 2.5, 7.0, 150, f)
    g["stokes_poly_force"] = dict(mesh=L(mesh), rho=2.5, pressure=7.0, G=150, lo=L(lo), di=L(di), up=L(up), rhs=L(rhs),
                                  solution=L(o.thomas(lo, di, up, rhs)))

    # ---- C2 static diffusion, irregular mesh (src/bin/many_divisions_irregular_mesh.rs:6-13)
    mesh = np.array(ing["1dbar_many_divisions_irregular.obj"]["nodes"])
    lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150)
    g["C2"] = dict(mesh=L(mesh), mu=1.0, b=1.0, bc=[1.0, 15.0], G=150, lo=L(lo), di=L(di), up=L(up), rhs=L(rhs),
                   solution=L(o.thomas(lo, di, up, rhs)))

    # ---- C3 time dependent, 10^4 steps on the big_mesh.obj node set
    mesh = big_mesh_nodes(os.path.join(adir, "big_mesh.obj"))
    bands = o.td_assemble(mesh, 1.0, 1.0, 150)
    st0 = o.td_initial_state(np.zeros(len(mesh) - 2), len(mesh), (1.0, 15.0))
    snaps = {}
    st = st0
    done = 0
    for upto in (1, 10, 100, 1000, 10000):
        st = o.td_run(bands, st, 1e-3, (1.0, 15.0), upto - done)
        done = upto
        snaps[str(upto)] = L(st)
    g["C3"] = dict(mesh=L(mesh), mu=1.0, b=1.0, bc=[1.0, 15.0], G=150, dt=1e-3, bands=[L(a) for a in bands], states=snaps)

    # ---- the reference's own unit-test inputs
    unit = {}
    for name, mesh in [("3p", [0.0, 0.5, 1.0]), ("4p", [0.0, 0.33, 0.66, 1.0]), ("5p", [0.0, 0.25, 0.5, 0.75, 1.0])]:
        lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (0.0, 1.0), 150)
        unit["ti_" + name] = dict(mesh=mesh, lo=L(lo), di=L(di), up=L(up), rhs=L(rhs), solution=L(o.thomas(lo, di, up, rhs)))
    bands = o.td_assemble([0.0, 0.5, 1.0], 1.0, 1.0, 150)
    st = o.td_run(bands, o.td_initial_state([15.0], 3, (1.0, 0.0)), 0.01, (1.0, 0.0), 1000)
    unit["td_3p"] = dict(mesh=[0.0, 0.5, 1.0], bands=[L(a) for a in bands], state_1000=L(st))
    lo, di, up, rhs = o.stokes_assemble([0.0, 0.333, 0.666, 1.0], 1.0, 1.0, 150, 10.0)
    unit["stokes_4p"] = dict(mesh=[0.0, 0.333, 0.666, 1.0], lo=L(lo), di=L(di), up=L(up), rhs=L(rhs),
                             solution=L(o.thomas(lo, di, up, rhs)))
    g["unit"] = unit

    # ---- a slice of the C5 synthetic ensemble (4 bars x 64 nodes, and 2 bars x 1024 nodes), 50 steps
    ens = {}
    for tag, ids, n, steps in [("small", [0, 1, 255, 65535], 64, 50), ("full_n", [0, 32767], 1024, 20)]:
        meshes = synthetic.ensemble_meshes(ids, n)
        ic = synthetic.ensemble_initial_conditions(meshes, ids)
        bands = o.ensemble_td_assemble(meshes, 1.0, 1.0, 150)
        st = np.stack([o.td_initial_state(ic[k], n, (1.0, 15.0)) for k in range(len(ids))])
        h = 1.0 / (n - 1)
        dt = 1e-8 if n == 1024 else 0.05 * h * h
        out = o.ensemble_td_run(bands, st, dt, (1.0, 15.0), steps)
        ens[tag] = dict(ids=ids, n=n, steps=steps, dt=dt, mesh_row0=L(meshes[0]), state=L(out),
                        mdi_row0=L(bands[1][0]), sdi_row0=L(bands[4][0]))
    g["ensemble"] = ens

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f)
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
