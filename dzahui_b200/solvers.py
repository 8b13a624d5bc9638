"""Solvers behind the reference's `DiffEquationSolver` trait (solvers/solver_trait.rs:12-23):

    XSolver.new(params, mesh, gauss_step)   -> time_independent.rs:57, time_dependent.rs:62, dim1.rs:77
    solver.solve(time_step) -> list of n nodal values (a freshly owned vector, like the Rust Vec<f64>)

`mesh` may be a `Mesh`, or any sequence of sorted node coordinates (the reference takes `Vec<f64>`).
Every call goes through the C ABI (include/dzahui_b200.h) to the CUDA kernels; nothing is computed on the CPU."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check
from .mesh import Mesh
from .params import DiffussionParamsTimeDependent, DiffussionParamsTimeIndependent, StokesParams1D

ASSEMBLY_AUTO, ASSEMBLY_FAITHFUL, ASSEMBLY_FAST = 0, 1, 2
SOLVE_AUTO, SOLVE_FAITHFUL, SOLVE_PARALLEL = 0, 1, 2


def make_options(assembly_mode=ASSEMBLY_AUTO, solve_mode=SOLVE_AUTO, refine_steps=-1, stored_operators=False):
    """stored_operators: ensembles step from pre-scaled operator arrays (56 B per DOF-step) instead of rebuilding the
    operator on chip (40 B)"""
    return _lib.Options(int(assembly_mode), int(solve_mode), int(refine_steps), 1 if stored_operators else 0)


def _as_mesh(mesh):
    return mesh if isinstance(mesh, Mesh) else Mesh.from_nodes(mesh)


class TriDiag:
    """Read-only view that lets tests index bands like the reference's dense `Array2`: K[i, j]."""

    def __init__(self, lo, di, up):
        self.lo, self.di, self.up = lo, di, up
        self.n = len(di)

    def __getitem__(self, ij):
        i, j = ij
        if not (0 <= i < self.n and 0 <= j < self.n):
            raise IndexError(ij)
        if j == i:
            return float(self.di[i])
        if j == i - 1:
            return float(self.lo[i])
        if j == i + 1:
            return float(self.up[i])
        return 0.0


class DiffEquationSolver:
    """trait DiffEquationSolver: Debug  { fn solve(&mut self, time_step: f64) -> Result<Vec<f64>, Error>; }"""

    def __init__(self):
        self._h = None
        self._lib = _lib.load()
        self.n = 0
        self._mesh = None

    def _finish(self, handle, mesh):
        self._h = handle
        self._mesh = mesh  # keeps the device mesh alive
        n = C.c_size_t()
        check(self._lib.dzb_solver_node_count(self._h, C.byref(n)), self._lib)
        self.n = n.value

    def solve(self, time_step=0.0):
        out = np.empty(self.n, dtype=np.float64)
        check(self._lib.dzb_solver_solve(self._h, float(time_step), out.ctypes.data_as(_lib.c_double_p), self.n), self._lib)
        return out

    def rebuild(self, mesh=None):
        """XSolver::new again into the same device buffers (same node count and params): re-assemble for `mesh`"""
        m = self._mesh if mesh is None else _as_mesh(mesh)
        check(self._lib.dzb_solver_rebuild(self._h, m._h), self._lib)
        self._mesh = m

    def solve_into(self, time_step, host_ptr):
        """solve() into a caller-owned (e.g. pinned) host buffer of n doubles"""
        check(self._lib.dzb_solver_solve(self._h, float(time_step), C.cast(C.c_void_p(int(host_ptr)), _lib.c_double_p), self.n),
              self._lib)

    def solve_device(self, time_step=0.0):
        """Same step, result left on the device; returns the raw device pointer (valid until the next call)."""
        p = C.c_void_p()
        check(self._lib.dzb_solver_solve_device(self._h, float(time_step), C.byref(p)), self._lib)
        return p.value

    def path(self):
        """which kernels the last solve ran through"""
        buf = C.create_string_buffer(128)
        check(self._lib.dzb_solver_path(self._h, buf, 128), self._lib)
        return buf.value.decode()

    def _bands(self, which, with_rhs):
        lo, di, up = (np.empty(self.n) for _ in range(3))
        rhs = np.empty(self.n) if with_rhs else None
        check(self._lib.dzb_solver_bands(self._h, which, lo.ctypes.data_as(_lib.c_double_p), di.ctypes.data_as(_lib.c_double_p),
                                         up.ctypes.data_as(_lib.c_double_p),
                                         rhs.ctypes.data_as(_lib.c_double_p) if with_rhs else None, self.n), self._lib)
        return lo, di, up, rhs

    def close(self):
        if self._h:
            self._lib.dzb_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _StaticSolver(DiffEquationSolver):
    @property
    def stiffness_matrix(self):
        lo, di, up, _ = self._bands(0, True)
        return TriDiag(lo, di, up)

    @property
    def b_vector(self):
        return self._bands(0, True)[3]

    def bands(self):
        """(lo, di, up, rhs) numpy arrays: lo[i] = K[i][i-1], up[i] = K[i][i+1]"""
        return self._bands(0, True)


class DiffussionSolverTimeIndependent(_StaticSolver):
    """-mu u_xx + b u_x = 0, Dirichlet both ends (diffusion_solver/time_independent.rs:46-53)."""

    def __init__(self, params: DiffussionParamsTimeIndependent, mesh, gauss_step: int, options=None):
        super().__init__()
        m = _as_mesh(mesh)
        h = C.c_void_p()
        check(self._lib.dzb_diffusion_ti_create(m._h, params.mu, params.b, params.boundary_conditions[0],
                                                params.boundary_conditions[1], int(gauss_step),
                                                C.byref(options) if options is not None else None, C.byref(h)), self._lib)
        self._finish(h, m)
        self.boundary_conditions = tuple(params.boundary_conditions)
        self.gauss_step = int(gauss_step)
        self.mu, self.b = params.mu, params.b

    new = classmethod(lambda cls, params, mesh, gauss_step, options=None: cls(params, mesh, gauss_step, options))

    def __repr__(self):
        return (f"DiffussionSolverTimeIndependent {{ boundary_conditions: {list(self.boundary_conditions)}, gauss_step: "
                f"{self.gauss_step}, mu: {self.mu}, b: {self.b}, n: {self.n} }}")


class StokesSolver1D(_StaticSolver):
    """(1/rho) p_x = f with a hydrostatic-pressure Dirichlet row (stokes_solver/dim1.rs:66-72)."""

    def __init__(self, params: StokesParams1D, mesh, gauss_step: int, options=None):
        super().__init__()
        m = _as_mesh(mesh)
        h = C.c_void_p()
        f = params.force_function
        cb = _lib.FORCE_FN(lambda x, _ctx: float(f(x)))  # only called during construction, like the borrowed closure
        check(self._lib.dzb_stokes1d_create(m._h, params.rho, params.hydrostatic_pressure, cb, None, int(gauss_step),
                                            C.byref(options) if options is not None else None, C.byref(h)), self._lib)
        self._finish(h, m)
        self.gauss_step = int(gauss_step)
        self.hydrostatic_pressure = params.hydrostatic_pressure
        self.rho = params.rho

    new = classmethod(lambda cls, params, mesh, gauss_step, options=None: cls(params, mesh, gauss_step, options))

    def __repr__(self):
        return (f"StokesSolver1D {{ gauss_step: {self.gauss_step}, hydrostatic_pressure: {self.hydrostatic_pressure}, "
                f"rho: {self.rho}, n: {self.n} }}")


StaticPressureSolver = StokesSolver1D  # stokes_solver/mod.rs:9


class DiffussionSolverTimeDependent(DiffEquationSolver):
    """u_t - mu u_xx + b u_x = 0, explicit Euler with a consistent mass matrix (time_dependent.rs:49-58)."""

    def __init__(self, params: DiffussionParamsTimeDependent, mesh, integration_step: int, options=None):
        super().__init__()
        m = _as_mesh(mesh)
        ic = np.ascontiguousarray(np.asarray(params.initial_conditions, dtype=np.float64))
        h = C.c_void_p()
        check(self._lib.dzb_diffusion_td_create(m._h, params.mu, params.b, params.boundary_conditions[0],
                                                params.boundary_conditions[1],
                                                ic.ctypes.data_as(_lib.c_double_p) if ic.size else None, ic.size,
                                                int(integration_step), C.byref(options) if options is not None else None,
                                                C.byref(h)), self._lib)
        self._finish(h, m)
        self.boundary_conditions = tuple(params.boundary_conditions)
        self.initial_conditions = list(params.initial_conditions)
        self.integration_step = int(integration_step)
        self.mu, self.b = params.mu, params.b

    new = classmethod(lambda cls, params, mesh, integration_step, options=None: cls(params, mesh, integration_step, options))

    @property
    def stiffness_matrix(self):
        lo, di, up, _ = self._bands(0, False)
        return TriDiag(lo, di, up)

    @property
    def mass_matrix(self):
        lo, di, up, _ = self._bands(1, False)
        return TriDiag(lo, di, up)

    def bands(self):
        """(mlo, mdi, mup, slo, sdi, sup)"""
        m = self._bands(1, False)
        s = self._bands(0, False)
        return m[0], m[1], m[2], s[0], s[1], s[2]

    @property
    def state(self):
        out = np.empty(self.n)
        check(self._lib.dzb_solver_state(self._h, out.ctypes.data_as(_lib.c_double_p), self.n), self._lib)
        return out

    def __repr__(self):
        return (f"DiffussionSolverTimeDependent {{ boundary_conditions: {list(self.boundary_conditions)}, integration_step: "
                f"{self.integration_step}, mu: {self.mu}, b: {self.b}, n: {self.n} }}")


class NoSolver(DiffEquationSolver):       # solvers/fem/mod.rs:36-42
    def __init__(self):
        self._h = None
        self.n = 0

    def solve(self, _time_step=0.0):
        return np.empty(0)

    def __repr__(self):
        return "NoSolver"


class Solver:
    """`enum Solver` (solvers/fem/mod.rs:27-33): an equation tagged with its params, turned into a boxed solver the
    way DzahuiWindow::run does (dzahui_window.rs:738-800)."""
    DiffussionSolverTimeIndependent = "DiffussionSolverTimeIndependent"
    DiffussionSolverTimeDependent = "DiffussionSolverTimeDependent"
    Stokes1DSolver = "Stokes1DSolver"
    Stokes2DSolver = "Stokes2DSolver"
    NONE = "None"

    def __init__(self, kind, params=None):
        self.kind, self.params = kind, params

    def instantiate(self, mesh, integration_iteration, options=None) -> DiffEquationSolver:
        if self.kind == Solver.DiffussionSolverTimeDependent:
            return DiffussionSolverTimeDependent(self.params, mesh, integration_iteration, options)
        if self.kind == Solver.DiffussionSolverTimeIndependent:
            return DiffussionSolverTimeIndependent(self.params, mesh, integration_iteration, options)
        if self.kind == Solver.Stokes1DSolver:
            return StokesSolver1D(self.params, mesh, integration_iteration, options)
        if self.kind == Solver.Stokes2DSolver:
            raise NotImplementedError("Not implemented yet!")  # dzahui_window.rs:792-794
        return NoSolver()
