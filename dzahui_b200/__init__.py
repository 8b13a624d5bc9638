"""dzahui_b200 — B200-native (sm_100a) 1D-FEM hot path of Dzahui behind the reference's solver API.

Public names follow the reference crate root (src/lib/mod.rs:8-14, solvers/mod.rs:2-10):
StokesParams, DiffussionParams, the three solvers, DiffEquationSolver, Solver, plus Mesh ingestion and the
ensemble driver.  All compute is hand-written CUDA reached through the C ABI in include/dzahui_b200.h."""
from . import _lib
from .ensemble import EnsembleTimeDependent
from .euler import EulerEnsemble, EulerSolver
from .errors import CudaError, Error, Integration, Invalid, Io, MeshParse, PieceWiseDims, WrongDims
from .mesh import Mesh, Mesh2D, MeshBuilder
from .params import (BuilderPanic, DiffussionParams, DiffussionParamsTimeDependent, DiffussionParamsTimeIndependent,
                     StaticPressureParams, StaticPressureParamsBuilder, StokesParams, StokesParams1D, StokesParams1DBuilder)
from .writer import Writer, format_f64
from .solvers import (ASSEMBLY_AUTO, ASSEMBLY_FAITHFUL, ASSEMBLY_FAST, SOLVE_AUTO, SOLVE_FAITHFUL, SOLVE_PARALLEL,
                      DiffEquationSolver, DiffussionSolverTimeDependent, DiffussionSolverTimeIndependent, NoSolver, Solver,
                      StaticPressureSolver, StokesSolver1D, TriDiag, make_options)


def quad_pair(n, k):
    """gauss_legendre::quad_pair(n, k) -> (theta, weight)   (quadrature/gauss_legendre.rs:5908-5927)"""
    import ctypes as C
    from .errors import check
    lib = _lib.load()
    th, w = C.c_double(), C.c_double()
    check(lib.dzb_quad_pair(n, k, C.byref(th), C.byref(w)), lib)
    return th.value, w.value


def quad_table(gauss_step):
    """(cos(theta_j), w_j) for j = 1..gauss_step-1, the nodes the reference's integration loops visit"""
    import numpy as np
    from .errors import check
    lib = _lib.load()
    m = max(int(gauss_step) - 1, 0)
    x, w = np.zeros(m), np.zeros(m)
    check(lib.dzb_quad_table(int(gauss_step), x.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p)), lib)
    return x, w


def device_count():
    import ctypes as C
    lib = _lib.load()
    c = C.c_int()
    lib.dzb_device_count(C.byref(c))
    return c.value


def set_device(device):
    from .errors import check
    lib = _lib.load()
    check(lib.dzb_set_device(int(device)), lib)


def set_stream(cuda_stream_ptr):
    lib = _lib.load()
    lib.dzb_set_stream(cuda_stream_ptr)


def synchronize():
    from .errors import check
    lib = _lib.load()
    check(lib.dzb_synchronize(), lib)


def launch_count(reset=False):
    import ctypes as C
    lib = _lib.load()
    c = C.c_ulonglong()
    lib.dzb_launch_count(C.byref(c), 1 if reset else 0)
    return c.value


class PinnedBuffer:
    """Page-locked host memory from dzb_host_alloc, viewed as a numpy array (`.array`): a result buffer for
    `solve(out=...)` / `solve_into(...)` that the D2H copy fills at PCIe speed.  Freed by close() / garbage collection."""

    def __init__(self, shape, dtype="float64"):
        import ctypes as C
        import numpy as np
        from .errors import check
        self._lib = _lib.load()
        dt = np.dtype(dtype)
        count = int(np.prod(shape))
        self._p = C.c_void_p()
        check(self._lib.dzb_host_alloc(C.byref(self._p), count * dt.itemsize), self._lib)
        self._raw = (C.c_char * (count * dt.itemsize)).from_address(self._p.value)
        self.array = np.frombuffer(self._raw, dtype=dt, count=count).reshape(shape)
        self.ptr = self._p.value

    def close(self):
        if self._p:
            self.array = None
            self._raw = None
            self._lib.dzb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
