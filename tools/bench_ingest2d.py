#!/usr/bin/env python3
"""Boundary-vertex detection on a large structured triangle grid: device hash table (dzb_boundary_vertices_device)
vs the CPU oracle's HashMap-style counting.  Prints one JSON line.  Algorithmic bytes: 12 B per triangle read (3 u32)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def grid_triangles(nx, ny):
    i, j = np.meshgrid(np.arange(nx - 1, dtype=np.uint32), np.arange(ny - 1, dtype=np.uint32), indexing="xy")
    v = (j * nx + i).ravel()
    t = np.empty((v.size * 2, 3), dtype=np.uint32)
    t[0::2] = np.stack([v, v + 1, v + nx], axis=1)
    t[1::2] = np.stack([v + 1, v + nx + 1, v + nx], axis=1)
    return t


def main():
    import torch
    import dzahui_b200 as dz
    from dzahui_b200 import _lib
    import oracle_lib as o
    nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
    dz.set_device(0)
    lib = _lib.load()
    tri = grid_triangles(nx, ny)
    rng = np.random.default_rng(0)
    tri = tri[rng.permutation(len(tri))]
    d_tri = torch.from_numpy(tri.astype(np.int32).ravel()).cuda()
    d_out = torch.empty(nx * ny, dtype=torch.int32, device="cuda")
    n_out = C.c_size_t()

    def run():
        rc = lib.dzb_boundary_vertices_device(C.c_void_p(d_tri.data_ptr()), len(tri), nx * ny, C.c_void_p(d_out.data_ptr()), C.byref(n_out))
        assert rc == 0, lib.dzb_last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / reps
    got = d_out[:n_out.value].cpu().numpy().astype(np.uint32)
    assert n_out.value == 2 * (nx + ny) - 4 and np.all(np.diff(got.astype(np.int64)) > 0)
    # CPU oracle on a bounded sample (a smaller grid), scaled per triangle
    sx = min(nx, 501)
    st = grid_triangles(sx, sx)
    t0 = time.perf_counter()
    want = o.boundary_vertices(st)
    cpu_s = time.perf_counter() - t0
    assert len(want) == 4 * sx - 4
    small = torch.from_numpy(st.astype(np.int32).ravel()).cuda()
    rc = lib.dzb_boundary_vertices_device(C.c_void_p(small.data_ptr()), len(st), sx * sx, C.c_void_p(d_out.data_ptr()), C.byref(n_out))
    assert rc == 0 and np.array_equal(d_out[:n_out.value].cpu().numpy().astype(np.uint32), want)
    print(json.dumps({"what": "boundary vertices of a triangle mesh", "triangles": len(tri), "vertices": nx * ny,
                      "gpu_ms": gpu_s * 1e3, "gpu_triangles_per_s": len(tri) / gpu_s,
                      "gpu_note": "includes cudaMalloc/cudaFree of the hash table and the result-count read-back",
                      "cpu_triangles_per_s": len(st) / cpu_s, "cpu_sample_triangles": len(st), "cpu_threads": 1,
                      "parity": "device list == oracle list on the sample; count and ordering checked at full size"}))


if __name__ == "__main__":
    main()
