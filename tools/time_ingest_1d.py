"""Times MeshBuilder::build_mesh_1d's replacement (dzb_mesh_ingest_obj_1d: text -> sorted nodes, GL vertices and indices on
the device) on a synthetic Wavefront file of n vertices.  Usage: python tools/time_ingest_1d.py [n ...]"""
import os, sys, time, json, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dzahui_b200 as dz


def main():
    out = {}
    for n in [int(float(a)) for a in sys.argv[1:]] or [1_000_000]:
        x = np.sort(np.random.default_rng(1).uniform(0.0, 100.0, n))
        path = os.path.join(tempfile.gettempdir(), f"bar_{n}.obj")
        with open(path, "w") as f:
            f.write("# synthetic bar\n")
            f.write("".join(f"v {v!r} 0.0 0.0\n" for v in x.tolist()))
        size = os.path.getsize(path)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            m = dz.Mesh.builder(path).build_mesh_1d()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert m.n == n
        out[n] = {"seconds": best, "file_MB": size / 1e6, "MB_per_s": size / 1e6 / best, "vertices_per_s": n / best}
        os.remove(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
