"""Host mirror of the reference's EulerSolver (solvers/euler/mod.rs:26-59) over the C ABI.

The reference steps one trajectory with a Rust closure; a host closure cannot run on the device, so the device version
takes the derivative function either by its affine coefficients, f(values) = c[0] + c[1]*values[0] + ... (which covers every
closure of the reference's integration tests, tests/euler.rs), or — for any closed-form expression of the old values and
per-trajectory parameters — as an expression (`EulerEnsemble.from_expression`, compiled to a postfix program the device
interprets: dzb_euler_create_program), and steps many trajectories at once."""
import ctypes as C
import math

import numpy as np

from . import _lib
from .errors import check


# ---------------------------------------------------------------------------------------------- derivative expressions
OPCODES = {"const": 0, "value": 1, "param": 2, "add": 3, "sub": 4, "mul": 5, "div": 6, "neg": 7, "sqrt": 8, "abs": 9, "sin": 10,
           "cos": 11, "exp": 12, "log": 13}
_HOST_UNARY = {"neg": lambda a: -a, "sqrt": math.sqrt, "abs": abs, "sin": math.sin, "cos": math.cos, "exp": math.exp, "log": math.log}


class Expr:
    """A derivative function under construction: Python arithmetic on `Expr` objects records a postfix program
    [(opcode, arg, value), ...] — what the reference's closure body would compute, operation for operation."""

    def __init__(self, rpn):
        self.rpn = list(rpn)

    @staticmethod
    def wrap(x):
        return x if isinstance(x, Expr) else Expr([("const", 0, float(x))])

    @staticmethod
    def value(j):
        return Expr([("value", int(j), 0.0)])

    @staticmethod
    def param(j):
        return Expr([("param", int(j), 0.0)])

    def _bin(self, other, op, swap=False):
        a, b = (Expr.wrap(other), self) if swap else (self, Expr.wrap(other))
        return Expr(a.rpn + b.rpn + [(op, 0, 0.0)])

    def _un(self, op):
        return Expr(self.rpn + [(op, 0, 0.0)])

    __add__ = lambda self, o: self._bin(o, "add")
    __radd__ = lambda self, o: self._bin(o, "add", True)
    __sub__ = lambda self, o: self._bin(o, "sub")
    __rsub__ = lambda self, o: self._bin(o, "sub", True)
    __mul__ = lambda self, o: self._bin(o, "mul")
    __rmul__ = lambda self, o: self._bin(o, "mul", True)
    __truediv__ = lambda self, o: self._bin(o, "div")
    __rtruediv__ = lambda self, o: self._bin(o, "div", True)
    __neg__ = lambda self: self._un("neg")
    __abs__ = lambda self: self._un("abs")

    def eval(self, values, params=()):
        """the same program on the host in Python floats (IEEE double, same order): what the Rust closure returns"""
        st = []
        for op, arg, val in self.rpn:
            if op == "const":
                st.append(val)
            elif op == "value":
                st.append(float(values[arg]))
            elif op == "param":
                st.append(float(params[arg]))
            elif op in _HOST_UNARY:
                st.append(_HOST_UNARY[op](st.pop()))
            else:
                b, a = st.pop(), st.pop()
                st.append(a + b if op == "add" else a - b if op == "sub" else a * b if op == "mul" else a / b)
        return st[0]


def sqrt(x):
    return Expr.wrap(x)._un("sqrt")


def sin(x):
    return Expr.wrap(x)._un("sin")


def cos(x):
    return Expr.wrap(x)._un("cos")


def exp(x):
    return Expr.wrap(x)._un("exp")


def log(x):
    return Expr.wrap(x)._un("log")


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

    @classmethod
    def from_expression(cls, f, values, params=None):
        """`f(v, p)` builds the derivative function from v[j] (the old values [y^(n-1), ..., y, t]) and p[j] (this
        trajectory's parameters) with + - * / abs() and this module's sqrt / sin / cos / exp / log, e.g. the pendulum
        `lambda v, p: -p[0] * sin(v[1])` on values [omega, theta, t].  One program for all trajectories."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        if v.ndim != 2:
            raise ValueError("values must be [n_traj][width]")
        self.n_traj, self.width = v.shape
        par = np.zeros((self.n_traj, 0)) if params is None else np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        if par.ndim != 2 or par.shape[0] != self.n_traj:
            raise ValueError("params must be [n_traj][n_params]")
        self.expr = Expr.wrap(f([Expr.value(j) for j in range(self.width)], [Expr.param(j) for j in range(par.shape[1])]))
        ops = (_lib.EulerOp * len(self.expr.rpn))(*[_lib.EulerOp(OPCODES[op], arg, val) for op, arg, val in self.expr.rpn])
        h = C.c_void_p()
        check(self._lib.dzb_euler_create_program(self.n_traj, self.width, C.cast(ops, C.c_void_p), len(self.expr.rpn),
                                                 par.ctypes.data_as(_lib.c_double_p), par.shape[1], v.ctypes.data_as(_lib.c_double_p),
                                                 C.byref(h)), self._lib)
        self._h = h
        return self

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
        self._f = None

    @classmethod
    def new(cls, f, width):
        """EulerSolver::new(|val: &[f64; width]| ...) with the closure body written over `val[j]` (see EulerEnsemble.from_expression):
        `EulerSolver.new(lambda val: -9.81 * sin(val[1]), 3)`"""
        self = cls([0.0] * (int(width) + 1))
        self._f = f
        return self

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
            self._ens = (EulerEnsemble([self.coeffs], [v]) if self._f is None else
                         EulerEnsemble.from_expression(lambda val, _p: self._f(val), [v]))
        elif self._last != v:
            self._ens.values = [v]
        out = self._ens.do_step(step)[0]
        self._last = [float(x) for x in out]
        return list(self._last)
