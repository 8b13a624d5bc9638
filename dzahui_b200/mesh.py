"""Mesh ingestion, device resident: mirrors `Mesh::builder(path).build_mesh_1d(height_multiplier)`
(mesh/mod.rs:46-52, mesh/mesh_builder.rs:204-328) and `Mesh::filter_for_solving_1d` (mesh/mod.rs:54-68)."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check


class Mesh:
    """A 1D bar on the device: sorted node coordinates, GL vertices [x,0,0,r,g,b] x 2n, triangle indices and the
    element-to-node table.  Build with `Mesh.builder(path).build_mesh_1d(mult)` or `Mesh.from_nodes(x)`."""

    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib
        n = C.c_size_t()
        check(lib.dzb_mesh_node_count(handle, C.byref(n)), lib)
        self.n = n.value
        ml, pw = C.c_double(), C.c_double()
        tr = (C.c_float * 3)()
        check(lib.dzb_mesh_info(handle, C.byref(ml), C.byref(pw), tr), lib)
        self.max_length = ml.value          # Mesh.max_length  (mesh/mod.rs:31)
        self.prom_width = pw.value
        self.model_translation = np.array(list(tr), dtype=np.float32)  # translation part of Mesh.model_matrix

    @staticmethod
    def builder(location):
        return MeshBuilder(location)

    @classmethod
    def from_nodes(cls, nodes):
        """`mesh: Vec<f64>` as handed to XSolver::new — sorted ascending node coordinates."""
        lib = _lib.load()
        x = np.ascontiguousarray(np.asarray(nodes, dtype=np.float64))
        h = C.c_void_p()
        check(lib.dzb_mesh_from_nodes(x.ctypes.data_as(_lib.c_double_p), x.size, C.byref(h)), lib)
        return cls(h, lib)

    def set_nodes(self, nodes):
        """re-point a from_nodes mesh at new coordinates (same count) without reallocating device memory"""
        x = np.ascontiguousarray(np.asarray(nodes, dtype=np.float64))
        check(self._lib.dzb_mesh_set_nodes(self._h, x.ctypes.data_as(_lib.c_double_p), x.size), self._lib)
        self.max_length = float(-x[0] + x[-1])

    def filter_for_solving_1d(self):
        out = np.empty(self.n, dtype=np.float64)
        check(self._lib.dzb_mesh_nodes(self._h, out.ctypes.data_as(_lib.c_double_p), self.n), self._lib)
        return out

    @property
    def vertices(self):
        out = np.empty(12 * self.n, dtype=np.float64)
        check(self._lib.dzb_mesh_vertices(self._h, out.ctypes.data_as(_lib.c_double_p), out.size), self._lib)
        return out

    @property
    def indices(self):
        out = np.empty(6 * self.n - 6, dtype=np.uint32)
        check(self._lib.dzb_mesh_indices(self._h, out.ctypes.data_as(_lib.c_u32_p), out.size), self._lib)
        return out

    @property
    def elements(self):
        """element-to-node table, shape (n-1, 2): element e = (node e, node e+1)"""
        out = np.empty(2 * (self.n - 1), dtype=np.uint32)
        check(self._lib.dzb_mesh_elements(self._h, out.ctypes.data_as(_lib.c_u32_p), out.size), self._lib)
        return out.reshape(-1, 2)

    def device_nodes_ptr(self):
        p = C.c_void_p()
        check(self._lib.dzb_mesh_device_nodes(self._h, C.byref(p)), self._lib)
        return p.value

    def update_gradient_1d(self, device_solution_ptr):
        """Mesh::update_gradient_1d(|u|) + Drawable::get_vertices: returns the f32[12n] vertex buffer to upload."""
        out = np.empty(12 * self.n, dtype=np.float32)
        check(self._lib.dzb_mesh_update_gradient_1d(self._h, C.c_void_p(device_solution_ptr), self.n,
                                                    out.ctypes.data_as(_lib.c_float_p), out.size), self._lib)
        return out

    def close(self):
        if self._h:
            self._lib.dzb_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mesh2D:
    """A planar triangle mesh on the device (`MeshBuilder::build_mesh_2d`, mesh_builder.rs:342-500): GL vertices
    [a, b, 0, r, g, b] per vertex, triangle indices, and the sorted boundary vertices."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib
        nv, nt, nb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib.dzb_mesh2d_counts(handle, C.byref(nv), C.byref(nt), C.byref(nb)), lib)
        self.n_vertices, self.n_triangles, self.n_boundary = nv.value, nt.value, nb.value
        ml = C.c_double()
        tr = (C.c_float * 3)()
        check(lib.dzb_mesh2d_info(handle, C.byref(ml), tr), lib)
        self.max_length = ml.value
        self.model_translation = np.array(list(tr), dtype=np.float32)

    @property
    def vertices(self):
        out = np.empty(6 * self.n_vertices)
        check(self._lib.dzb_mesh2d_vertices(self._h, out.ctypes.data_as(_lib.c_double_p), out.size), self._lib)
        return out

    @property
    def indices(self):
        out = np.empty(3 * self.n_triangles, dtype=np.uint32)
        check(self._lib.dzb_mesh2d_indices(self._h, out.ctypes.data_as(_lib.c_u32_p), out.size), self._lib)
        return out

    @property
    def boundary_indices(self):
        out = np.empty(self.n_boundary, dtype=np.uint32)
        check(self._lib.dzb_mesh2d_boundary_indices(self._h, out.ctypes.data_as(_lib.c_u32_p), out.size), self._lib)
        return out

    def close(self):
        if self._h:
            self._lib.dzb_mesh2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MeshBuilder:
    def __init__(self, location):
        self.location = str(location)

    def build_mesh_2d(self):
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.dzb_mesh_ingest_obj_2d(self.location.encode(), C.byref(h)), lib)
        return Mesh2D(h, lib)

    def build_mesh_1d(self, height_multiplier=None):
        lib = _lib.load()
        h = C.c_void_p()
        has = 0 if height_multiplier is None else 1
        mult = 0.0 if height_multiplier is None else float(height_multiplier)
        check(lib.dzb_mesh_ingest_obj_1d(self.location.encode(), has, mult, C.byref(h)), lib)
        return Mesh(h, lib)
