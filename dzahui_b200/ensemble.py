"""Ensembles of independent time-dependent bars (BASELINE config 5): every bar is one
DiffussionSolverTimeDependent (time_dependent.rs:49-280) with its own mesh and initial conditions; one `step`
is one `solve(dt)` on every bar.  Bars shard across GPUs with no communication (one process per GPU)."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check


class EnsembleTimeDependent:
    def __init__(self, meshes, mu, b, boundary_conditions, initial_conditions, integration_step, options=None):
        self._lib = _lib.load()
        self._h = None
        meshes = np.ascontiguousarray(np.asarray(meshes, dtype=np.float64))
        if meshes.ndim != 2:
            raise ValueError("meshes must be [n_bars][n]")
        ic = np.ascontiguousarray(np.asarray(initial_conditions, dtype=np.float64))
        self.n_bars, self.n = meshes.shape
        n_ic = ic.shape[1] if ic.ndim == 2 else ic.size
        if ic.ndim == 2 and ic.shape[0] != self.n_bars:
            raise ValueError("initial_conditions must be [n_bars][n-2]")
        h = C.c_void_p()
        check(self._lib.dzb_ensemble_td_create(meshes.ctypes.data_as(_lib.c_double_p), self.n_bars, self.n, float(mu), float(b),
                                               float(boundary_conditions[0]), float(boundary_conditions[1]),
                                               ic.ctypes.data_as(_lib.c_double_p) if ic.size else None, n_ic,
                                               int(integration_step), C.byref(options) if options is not None else None,
                                               C.byref(h)), self._lib)
        self._h = h
        self.boundary_conditions = tuple(boundary_conditions)
        self.mu, self.b, self.integration_step = mu, b, int(integration_step)

    def step(self, time_step, steps=1):
        """`steps` x solve(time_step) on every bar; states stay on the device."""
        check(self._lib.dzb_ensemble_step(self._h, float(time_step), int(steps)), self._lib)

    def solve(self, time_step, out=None):
        """One solve(time_step) on every bar; returns all states [n_bars][n] on the host (out may be a pinned buffer)."""
        if out is None:
            out = np.empty((self.n_bars, self.n), dtype=np.float64)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else int(out)
        check(self._lib.dzb_ensemble_solve(self._h, float(time_step), C.c_void_p(ptr), self.n_bars * self.n), self._lib)
        return out

    def solve_into(self, time_step, host_ptr):
        check(self._lib.dzb_ensemble_solve(self._h, float(time_step), C.c_void_p(int(host_ptr)), self.n_bars * self.n), self._lib)

    def solve_f32(self, time_step, out=None):
        """One solve(time_step) on every bar; the states come back cast to f32 on the device (half the PCIe bytes): what the
        reference's renderer uploads (mesh/mod.rs:108-113).  `out`: float32 array [n_bars][n] or a raw (pinned) host pointer."""
        if out is None:
            out = np.empty((self.n_bars, self.n), dtype=np.float32)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else int(out)
        check(self._lib.dzb_ensemble_solve_f32(self._h, float(time_step), C.c_void_p(ptr), self.n_bars * self.n), self._lib)
        return out

    def state(self, out=None):
        if out is None:
            out = np.empty((self.n_bars, self.n), dtype=np.float64)
        check(self._lib.dzb_ensemble_state(self._h, C.c_void_p(out.ctypes.data), self.n_bars * self.n), self._lib)
        return out

    def read_bars(self, bar_ids):
        ids = np.ascontiguousarray(np.asarray(bar_ids, dtype=np.uint64))
        out = np.empty((ids.size, self.n), dtype=np.float64)
        check(self._lib.dzb_ensemble_read_bars(self._h, ids.ctypes.data_as(_lib.c_size_p), ids.size,
                                               out.ctypes.data_as(_lib.c_double_p)), self._lib)
        return out

    def state_device_ptr(self):
        p = C.c_void_p()
        check(self._lib.dzb_ensemble_state_device(self._h, C.byref(p)), self._lib)
        return p.value

    def bands(self, bar):
        """(mlo, mdi, mup, slo, sdi, sup) of one bar"""
        res = []
        for which in (1, 0):
            lo, di, up = (np.empty(self.n) for _ in range(3))
            check(self._lib.dzb_ensemble_bands(self._h, int(bar), which, lo.ctypes.data_as(_lib.c_double_p),
                                               di.ctypes.data_as(_lib.c_double_p), up.ctypes.data_as(_lib.c_double_p), self.n),
                  self._lib)
            res += [lo, di, up]
        return tuple(res)

    def step_kernel(self):
        """name of the kernel one step dispatches to (final after the first step)"""
        buf = C.create_string_buffer(96)
        check(self._lib.dzb_ensemble_step_kernel(self._h, buf, 96), self._lib)
        return buf.value.decode()

    def close(self):
        if self._h:
            self._lib.dzb_ensemble_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
