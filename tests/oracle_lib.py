"""ctypes binding of the CPU oracle (oracle/libdzahui_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libdzahui_oracle.so")

OK, WRONG_DIMS, INTEGRATION, PIECEWISE_DIMS, MESH_PARSE, IO = range(6)

_dp = C.POINTER(C.c_double)
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dzo_basis_evaluate.restype = C.c_double
        _lib.dzo_basis_evaluate.argtypes = [_dp, C.c_size_t, C.c_size_t, C.c_double, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def _chk(rc):
    if rc != OK:
        raise OracleError(rc)


def hardware_threads():
    return lib().dzo_hardware_threads()


def quad_pair(n, k):
    th, w = C.c_double(), C.c_double()
    _chk(lib().dzo_quad_pair(C.c_size_t(n), C.c_size_t(k), C.byref(th), C.byref(w)))
    return th.value, w.value


def quad_table(G):
    x = np.zeros(max(G - 1, 0))
    w = np.zeros(max(G - 1, 0))
    _chk(lib().dzo_quad_table(C.c_size_t(G), _p(x), _p(w)))
    return x, w


def linear_basis(mesh):
    mesh = _arr(mesh)
    out = np.zeros((len(mesh), 12))
    _chk(lib().dzo_linear_basis(_p(mesh), C.c_size_t(len(mesh)), _p(out)))
    res = []
    for row in out:
        npoly = int(row[11])
        res.append((list(row[0:npoly]), list(row[4:4 + npoly]), list(row[8:8 + npoly - 1])))
    return res


def basis_evaluate(mesh, i, x, derivative=False):
    mesh = _arr(mesh)
    return lib().dzo_basis_evaluate(_p(mesh), len(mesh), i, x, int(derivative))


def ti_assemble(mesh, mu, b, bc, G, mode=0, threads=1):
    mesh = _arr(mesh)
    n = len(mesh)
    lo, di, up, rhs = (np.zeros(n) for _ in range(4))
    _chk(lib().dzo_ti_assemble(_p(mesh), C.c_size_t(n), C.c_double(mu), C.c_double(b), C.c_double(bc[0]),
                               C.c_double(bc[1]), C.c_size_t(G), mode, threads, _p(lo), _p(di), _p(up), _p(rhs)))
    return lo, di, up, rhs


def td_assemble(mesh, mu, b, G, mode=0, threads=1):
    mesh = _arr(mesh)
    n = len(mesh)
    a = [np.zeros(n) for _ in range(6)]
    _chk(lib().dzo_td_assemble(_p(mesh), C.c_size_t(n), C.c_double(mu), C.c_double(b), C.c_size_t(G), mode, threads,
                               *[_p(x) for x in a]))
    return tuple(a)  # mlo, mdi, mup, slo, sdi, sup


FORCE_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def stokes_assemble(mesh, rho, pressure, G, force, mode=0):
    """force: python callable x -> float, or a float for a constant force."""
    mesh = _arr(mesh)
    n = len(mesh)
    lo, di, up, rhs = (np.zeros(n) for _ in range(4))
    if callable(force):
        cb = FORCE_FN(lambda x, _ctx: float(force(x)))
        _chk(lib().dzo_stokes_assemble(_p(mesh), C.c_size_t(n), C.c_double(rho), C.c_double(pressure), C.c_size_t(G),
                                       cb, None, mode, _p(lo), _p(di), _p(up), _p(rhs)))
    else:
        _chk(lib().dzo_stokes_assemble_const_force(_p(mesh), C.c_size_t(n), C.c_double(rho), C.c_double(pressure),
                                                   C.c_size_t(G), C.c_double(force), mode, _p(lo), _p(di), _p(up),
                                                   _p(rhs)))
    return lo, di, up, rhs


def thomas(lo, di, up, rhs, kind="f64"):
    lo, di, up, rhs = map(_arr, (lo, di, up, rhs))
    n = len(rhs)
    out = np.zeros(n)
    fn = {"f64": lib().dzo_thomas, "f128": lib().dzo_thomas_f128, "f80": lib().dzo_thomas_f80}[kind]
    _chk(fn(_p(lo), _p(di), _p(up), _p(rhs), C.c_size_t(n), _p(out)))
    return out


def ti_intended_f128(mesh, mu, b, bc, G, brute=False, threads=1, bands=False):
    """Static diffusion system of the reference evaluated and solved in exact (__float128) arithmetic on its f64 inputs:
    -> u (f64), or (u, lo, di, up) with the exact bands rounded to f64.  brute: node-by-node sums instead of the moments."""
    mesh = _arr(mesh)
    n = len(mesh)
    u = np.zeros(n)
    lo, di, up = (np.zeros(n) for _ in range(3)) if bands else (None, None, None)
    null = C.cast(None, _dp)
    _chk(lib().dzo_ti_intended_f128(_p(mesh), C.c_size_t(n), C.c_double(mu), C.c_double(b), C.c_double(bc[0]), C.c_double(bc[1]),
                                    C.c_size_t(G), int(bool(brute)), int(threads), _p(u), _p(lo) if bands else null,
                                    _p(di) if bands else null, _p(up) if bands else null))
    return (u, lo, di, up) if bands else u


def td_intended_f128(mesh, mu, b, G, brute=False, threads=1):
    """M and S of the time-dependent solver (time_dependent.rs:110-247) evaluated in exact (__float128) arithmetic on the
    reference's f64 inputs, rounded to f64 at the end -> (mlo, mdi, mup, slo, sdi, sup)."""
    mesh = _arr(mesh)
    n = len(mesh)
    a = [np.zeros(n) for _ in range(6)]
    _chk(lib().dzo_td_intended_f128(_p(mesh), C.c_size_t(n), C.c_double(mu), C.c_double(b), C.c_size_t(G), int(bool(brute)),
                                    int(threads), *[_p(x) for x in a]))
    return tuple(a)


def thomas_dense(a, b):
    a = _arr(a)
    b = _arr(b)
    out = np.zeros(len(b))
    _chk(lib().dzo_thomas_dense(_p(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), _p(b), C.c_size_t(len(b)),
                                _p(out)))
    return out


def tridiag_matvec(lo, di, up, b, c):
    lo, di, up, b = map(_arr, (lo, di, up, b))
    out = np.zeros(len(b))
    _chk(lib().dzo_tridiag_matvec(_p(lo), _p(di), _p(up), _p(b), C.c_size_t(len(di)), C.c_size_t(len(b)),
                                  C.c_double(c), _p(out)))
    return out


def add(b, v):
    b, v = _arr(b), _arr(v)
    out = np.zeros(len(b))
    _chk(lib().dzo_add(_p(b), C.c_size_t(len(b)), _p(v), C.c_size_t(len(v)), _p(out)))
    return out


def td_initial_state(ic, n, bc):
    ic = _arr(ic)
    st = np.zeros(n)
    _chk(lib().dzo_td_initial_state(_p(ic), C.c_size_t(len(ic)), C.c_size_t(n), C.c_double(bc[0]), C.c_double(bc[1]),
                                    _p(st)))
    return st


def td_run(bands, state, dt, bc, steps, dense=False):
    """bands = (mlo, mdi, mup, slo, sdi, sup); returns the new state (input untouched)."""
    b = [_arr(x) for x in bands]
    st = _arr(state).copy()
    fn = lib().dzo_td_run_dense if dense else lib().dzo_td_run
    _chk(fn(*[_p(x) for x in b], _p(st), C.c_size_t(len(st)), C.c_double(dt), C.c_double(bc[0]), C.c_double(bc[1]),
            C.c_size_t(steps)))
    return st


def ensemble_td_assemble(meshes, mu, b, G, threads=1):
    meshes = _arr(meshes)
    B, n = meshes.shape
    a = [np.zeros((B, n)) for _ in range(6)]
    _chk(lib().dzo_ensemble_td_assemble(_p(meshes), C.c_size_t(B), C.c_size_t(n), C.c_double(mu), C.c_double(b),
                                        C.c_size_t(G), threads, *[_p(x) for x in a]))
    return tuple(a)


def ensemble_td_run(bands, state, dt, bc, steps, threads=1):
    b = [_arr(x) for x in bands]
    st = _arr(state).copy()
    B, n = st.shape
    _chk(lib().dzo_ensemble_td_run(*[_p(x) for x in b], _p(st), C.c_size_t(B), C.c_size_t(n), C.c_double(dt),
                                   C.c_double(bc[0]), C.c_double(bc[1]), C.c_size_t(steps), threads))
    return st


class EnsembleSession:
    """Resident CPU ensemble for timed legs: matrices built once, states advanced in place (no per-call copies)."""

    def __init__(self, bands, state, bc):
        b = [_arr(x) for x in bands]
        st = _arr(state)
        self.B, self.n = st.shape
        L = lib()
        L.dzo_ensemble_td_open.restype = C.c_void_p
        L.dzo_ensemble_td_advance.argtypes = [C.c_void_p, C.c_double, C.c_size_t, C.c_int]
        L.dzo_ensemble_td_read.argtypes = [C.c_void_p, _dp]
        L.dzo_ensemble_td_close.argtypes = [C.c_void_p]
        L.dzo_ensemble_td_close.restype = None
        self._h = L.dzo_ensemble_td_open(*[_p(x) for x in b], _p(st), C.c_size_t(self.B), C.c_size_t(self.n),
                                         C.c_double(bc[0]), C.c_double(bc[1]))

    def advance(self, dt, steps, threads=1):
        _chk(lib().dzo_ensemble_td_advance(self._h, float(dt), int(steps), int(threads)))

    def state(self):
        out = np.empty((self.B, self.n))
        _chk(lib().dzo_ensemble_td_read(self._h, _p(out)))
        return out

    def close(self):
        if self._h:
            lib().dzo_ensemble_td_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mesh_ingest_1d(path, height_multiplier=None):
    n = C.c_size_t(0)
    has = 0 if height_multiplier is None else 1
    mult = C.c_double(0.0 if height_multiplier is None else height_multiplier)
    _chk(lib().dzo_mesh_ingest_1d(path.encode(), has, mult, C.byref(n), None, C.c_size_t(0), None, C.c_size_t(0),
                                  None, None, None))
    nn = n.value
    v = np.zeros(12 * nn)
    idx = np.zeros(max(6 * nn - 6, 6), dtype=np.uint32)
    ml, pw = C.c_double(), C.c_double()
    tr = (C.c_float * 3)()
    _chk(lib().dzo_mesh_ingest_1d(path.encode(), has, mult, C.byref(n), _p(v), C.c_size_t(len(v)),
                                  idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(len(idx)), C.byref(ml),
                                  C.byref(pw), tr))
    return dict(n=nn, vertices=v, indices=idx[:max(6 * nn - 6, 0)] if nn > 1 else idx[:6], max_length=ml.value,
                prom_width=pw.value, translation=np.array(list(tr), dtype=np.float32))


def filter_for_solving_1d(vertices):
    vertices = _arr(vertices)
    out = np.zeros(len(vertices) // 12)
    k = lib().dzo_filter_for_solving_1d(_p(vertices), C.c_size_t(len(vertices)), _p(out))
    return out[:k]


def update_gradient_1d(vertices, velocity_norm):
    v = _arr(vertices).copy()
    s = _arr(velocity_norm)
    lib().dzo_update_gradient_1d(_p(v), C.c_size_t(len(v)), _p(s), C.c_size_t(len(s)))
    return v


def mesh_ingest_2d(path):
    """MeshBuilder::build_mesh_2d (mesh_builder.rs:342-500) -> dict(vertices, indices, boundary_indices, max_length, translation)"""
    nv, nt, nb = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    _chk(lib().dzo_mesh_ingest_2d(path.encode(), C.byref(nv), C.byref(nt), C.byref(nb), None, None, None, None, None))
    v = np.zeros(6 * nv.value)
    idx = np.zeros(3 * nt.value, dtype=np.uint32)
    bnd = np.zeros(nb.value, dtype=np.uint32)
    ml = C.c_double()
    tr = (C.c_float * 3)()
    u32p = C.POINTER(C.c_uint32)
    _chk(lib().dzo_mesh_ingest_2d(path.encode(), C.byref(nv), C.byref(nt), C.byref(nb), _p(v), idx.ctypes.data_as(u32p),
                                  bnd.ctypes.data_as(u32p), C.byref(ml), tr))
    return dict(n_vertices=nv.value, n_triangles=nt.value, vertices=v, indices=idx, boundary_indices=bnd, max_length=ml.value,
                translation=np.array(list(tr), dtype=np.float32))


def boundary_vertices(triangles):
    """the reference's edge-count boundary detection on an in-memory triangle list [T][3]"""
    t = np.ascontiguousarray(np.asarray(triangles, dtype=np.uint32)).reshape(-1, 3)
    out = np.zeros(int(t.max()) + 1 if t.size else 0, dtype=np.uint32)
    n = C.c_size_t(0)
    u32p = C.POINTER(C.c_uint32)
    _chk(lib().dzo_boundary_vertices(t.ctypes.data_as(u32p), C.c_size_t(t.shape[0]), out.ctypes.data_as(u32p), C.c_size_t(out.size),
                                     C.byref(n)))
    return out[:n.value].copy()


_EULER_FN = C.CFUNCTYPE(C.c_double, _dp, C.c_void_p)


def euler_do_step(values, step, f):
    """EulerSolver::new(f).do_step(values, step) (solvers/euler/mod.rs:37-59); f takes the list of old values"""
    v = _arr(values)
    out = np.empty_like(v)
    cb = _EULER_FN(lambda p, _ctx: float(f([p[j] for j in range(v.size)])))
    _chk(lib().dzo_euler_do_step(_p(v), C.c_size_t(v.size), C.c_double(step), cb, None, _p(out)))
    return out


def euler_affine_run(coeffs, values, step, steps):
    """`steps` do_step calls per trajectory with the closure c[0] + c[1]*v[0] + ...; returns the new [n_traj][width] array"""
    c, v = _arr(coeffs), _arr(values).copy()
    n_traj, width = v.shape
    assert c.shape == (n_traj, width + 1)
    _chk(lib().dzo_euler_affine_run(_p(c), _p(v), C.c_size_t(n_traj), C.c_size_t(width), C.c_double(step), C.c_size_t(steps)))
    return v
