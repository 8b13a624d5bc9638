"""Host mirror of the reference's EulerSolver (solvers/euler/mod.rs:26-59) over the C ABI.

The reference steps one trajectory with a Rust closure; a host closure cannot run on the device, so the device version
takes derivative functions that are affine in the state, f(values) = c[0] + c[1]*values[0] + ... (which covers every
closure of the reference's integration tests, tests/euler.rs), and steps many trajectories at once."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check


class EulerEnsemble:
    """n_traj independent trajectories; values[i] = [y^(n-1), ..., y, t] (2..5 entries), coeffs[i] = [c0, c1, ..., c_width]"""

    def __init__(self, coeffs, values):
        self._lib = _lib.load()
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64))
        if v.ndim != 2 or c.ndim != 2 or c.shape != (v.shape[0], v.shape[1] + 1):
            raise ValueError("values must be [n_traj][width] and coeffs [n_traj][width + 1]")
        self.n_traj, self.width = v.shape
        h = C.c_void_p()
        check(self._lib.dzb_euler_create(self.n_traj, self.width, c.ctypes.data_as(_lib.c_double_p), v.ctypes.data_as(_lib.c_double_p),
                                         C.byref(h)), self._lib)
        self._h = h

    def do_step(self, step, steps=1):
        """`steps` x EulerSolver::do_step(values, step) on every trajectory; returns the new [n_traj][width] values"""
        check(self._lib.dzb_euler_step(self._h, float(step), int(steps)), self._lib)
        return self.values

    def step_device(self, step, steps=1):
        """same, the states stay on the device (read them with `.values` when needed)"""
        check(self._lib.dzb_euler_step(self._h, float(step), int(steps)), self._lib)

    @property
    def values(self):
        out = np.empty((self.n_traj, self.width), dtype=np.float64)
        check(self._lib.dzb_euler_values(self._h, out.ctypes.data_as(_lib.c_double_p), out.size), self._lib)
        return out

    @values.setter
    def values(self, v):
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        check(self._lib.dzb_euler_set_values(self._h, v.ctypes.data_as(_lib.c_double_p), v.size), self._lib)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dzb_euler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EulerSolver:
    """EulerSolver::new(f) / do_step(values, step) -> values for ONE trajectory, f given by its affine coefficients:
    `EulerSolver.affine([0.0, -lam, 0.0]).do_step([quantity, time], 0.5)` is the reference's
    `EulerSolver::new(|val: &[f64; 2]| -lam * val[0]).do_step([quantity, time], 0.5)`."""

    def __init__(self, coeffs):
        self.coeffs = [float(c) for c in coeffs]
        if not 3 <= len(self.coeffs) <= 6:
            raise ValueError("EulerSolver takes [f64; 2] .. [f64; 5]")
        self._ens = None
        self._last = None

    @classmethod
    def affine(cls, coeffs):
        return cls(coeffs)

    @classmethod
    def constant(cls, value, width):
        return cls([float(value)] + [0.0] * width)

    def do_step(self, values, step):
        v = [float(x) for x in values]
        if len(v) != len(self.coeffs) - 1:
            raise ValueError("values and coefficients disagree on the width")
        if self._ens is None:
            self._ens = EulerEnsemble([self.coeffs], [v])
        elif self._last != v:
            self._ens.values = [v]
        out = self._ens.do_step(step)[0]
        self._last = [float(x) for x in out]
        return list(self._last)
