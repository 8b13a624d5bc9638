"""EulerSolver (reference solvers/euler/mod.rs:26-59): the oracle's do_step against the reference's own integration tests
(tests/euler.rs — the only fixtures the reference has for it: windows, not exact values), and the device ensemble against
the oracle, bit for bit."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def gpu(dz):
    assert dz.device_count() >= 1, "GPU tests need a CUDA device (there is no CPU fallback)"
    dz.set_device(0)
    return dz


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_first_order_ode(oracle):
    """tests/euler.rs:4-16"""
    pos, time = 0.0, 0.0
    while time <= 10.0:
        pos, time = oracle.euler_do_step([pos, time], 0.01, lambda v: 1.0)
    assert 9.9 <= pos <= 10.1


def test_oracle_second_order_ode_gravity(oracle):
    """tests/euler.rs:18-33"""
    vel, pos, time = 0.0, 100.0, 0.0
    while time <= 10.0:
        vel, pos, time = oracle.euler_do_step([vel, pos, time], 0.01, lambda v: -9.81)
    assert -100.0 <= vel <= -97.0
    assert -400.0 <= pos <= -389.0


def test_oracle_radioactive_decay(oracle):
    """tests/euler.rs:35-53"""
    lam = math.log(2.0) / 1600.0
    quantity, time = 1000.0, -1600.0
    while time <= 0.0:
        quantity, time = oracle.euler_do_step([quantity, time], 0.5, lambda v: -lam * v[0])
    assert -0.5 <= time <= 0.5
    assert 490.0 <= quantity <= 510.0


def test_oracle_affine_closure_is_the_closure(oracle):
    """the affine runner steps exactly like do_step with the equivalent Python closure"""
    rng = np.random.default_rng(5)
    for width in (2, 3, 4, 5):
        c = rng.uniform(-2.0, 2.0, (1, width + 1))
        v = rng.uniform(-1.0, 1.0, (1, width))
        got = oracle.euler_affine_run(c, v, 0.01, 25)[0]
        w = list(v[0])
        for _ in range(25):
            w = list(oracle.euler_do_step(w, 0.01, lambda x: sum_left_to_right(c[0], x)))
        assert np.array(w).tobytes() == got.tobytes()


def sum_left_to_right(c, x):
    f = c[0]
    for j, xj in enumerate(x):
        f = f + c[j + 1] * xj
    return f


def test_oracle_rejects_widths_the_reference_has_no_impl_for(oracle):
    with pytest.raises(Exception):
        oracle.euler_do_step([1.0], 0.1, lambda v: 0.0)
    with pytest.raises(Exception):
        oracle.euler_do_step([1.0] * 6, 0.1, lambda v: 0.0)


# ------------------------------------------------------------------------------------------------ device (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("width", [2, 3, 4, 5])
@pytest.mark.parametrize("steps", [1, 7, 1000])
def test_gpu_ensemble_matches_oracle_bits(gpu, oracle, width, steps):
    rng = np.random.default_rng(100 * width + steps)
    n_traj = 10_007
    c = rng.uniform(-1.5, 1.5, (n_traj, width + 1))
    v = rng.uniform(-2.0, 2.0, (n_traj, width))
    ens = gpu.EulerEnsemble(c, v)
    got = ens.do_step(1e-3, steps)
    want = oracle.euler_affine_run(c, v, 1e-3, steps)
    assert got.tobytes() == want.tobytes()
    # a second call continues from the device-resident state
    got2 = ens.do_step(2e-3, 3)
    assert got2.tobytes() == oracle.euler_affine_run(c, want, 2e-3, 3).tobytes()


@pytest.mark.gpu
def test_gpu_reference_integration_tests(gpu):
    """tests/euler.rs through the device path, written like the reference's tests"""
    solver = gpu.EulerSolver.constant(1.0, 2)          # EulerSolver::new(|_: &[f64; 2]| 1.0)
    pos, time = 0.0, 0.0
    while time <= 10.0:
        pos, time = solver.do_step([pos, time], 0.01)
    assert 9.9 <= pos <= 10.1
    solver = gpu.EulerSolver.constant(-9.81, 3)        # EulerSolver::new(|_: &[f64; 3]| -9.81)
    vel, pos, time = 0.0, 100.0, 0.0
    while time <= 10.0:
        vel, pos, time = solver.do_step([vel, pos, time], 0.01)
    assert -100.0 <= vel <= -97.0 and -400.0 <= pos <= -389.0
    lam = math.log(2.0) / 1600.0
    solver = gpu.EulerSolver.affine([0.0, -lam, 0.0])  # EulerSolver::new(|val: &[f64; 2]| -lam * val[0])
    quantity, time = 1000.0, -1600.0
    while time <= 0.0:
        quantity, time = solver.do_step([quantity, time], 0.5)
    assert -0.5 <= time <= 0.5 and 490.0 <= quantity <= 510.0


@pytest.mark.gpu
def test_gpu_euler_errors(gpu):
    with pytest.raises(gpu.Invalid):
        gpu.EulerEnsemble(np.zeros((3, 2)), np.zeros((3, 1)))       # [f64; 1]: no FunctionArguments impl
    with pytest.raises(gpu.Invalid):
        gpu.EulerEnsemble(np.zeros((3, 7)), np.zeros((3, 6)))       # [f64; 6]
    ens = gpu.EulerEnsemble(np.zeros((3, 3)), np.zeros((3, 2)))
    with pytest.raises(gpu.WrongDims):
        ens.values = np.zeros((2, 2))


# ------------------------------------------------------------------------------------------------ derivative expressions
def test_expression_records_the_closure_body():
    """the host side of dzb_euler_create_program: Python arithmetic on Expr objects is the closure body, operation for
    operation, and Expr.eval replays it in IEEE double (what the oracle's callback computes)"""
    from dzahui_b200 import euler as E
    v, p = [E.Expr.value(j) for j in range(3)], [E.Expr.param(0)]
    e = -p[0] * E.sin(v[1]) - 0.25 * v[0] + 1.0 / (1.0 + v[1] * v[1])
    vals, par = [0.3, -1.2, 0.0], [9.81]
    want = -par[0] * math.sin(vals[1]) - 0.25 * vals[0] + 1.0 / (1.0 + vals[1] * vals[1])
    assert e.eval(vals, par) == want
    assert [op for op, _, _ in e.rpn].count("mul") == 3 and e.rpn[-1][0] == "add"


@pytest.mark.gpu
def test_program_matches_oracle_bit_for_bit_on_exact_operations(gpu, oracle):
    """non-affine derivative functions built from + - * / neg abs sqrt: the device interpreter and the oracle's do_step with
    the same closure (solvers/euler/mod.rs:37-59) agree to the bit — Van der Pol and logistic growth with per-trajectory
    parameters, widths 2 and 3"""
    from dzahui_b200 import euler as E
    rng = np.random.default_rng(11)
    T = 300
    # Van der Pol, values [x', x, t]: f = mu (1 - x^2) x' - x
    vals = np.column_stack([rng.normal(size=T), rng.normal(size=T), np.zeros(T)])
    par = rng.uniform(0.1, 3.0, (T, 1))
    f = lambda v, p: p[0] * (1.0 - v[1] * v[1]) * v[0] - v[1]
    ens = E.EulerEnsemble.from_expression(f, vals, par)
    got = ens.do_step(0.01, 200)
    for i in (0, 1, 17, T - 1):
        w = vals[i].copy()
        for _ in range(200):
            w = oracle.euler_do_step(w, 0.01, lambda q: ens.expr.eval(q, par[i]))
        assert got[i].tobytes() == w.tobytes()
    ens.close()
    # logistic growth with an absolute value and a square root thrown in, values [y, t]: f = r y (1 - y / K) - sqrt(|y|) / 7
    vals = np.column_stack([rng.uniform(0.5, 2.0, T), np.zeros(T)])
    par = np.column_stack([rng.uniform(0.5, 1.5, T), rng.uniform(5.0, 10.0, T)])
    g = lambda v, p: p[0] * v[0] * (1.0 - v[0] / p[1]) - E.sqrt(abs(v[0])) / 7.0
    ens = E.EulerEnsemble.from_expression(g, vals, par)
    got = ens.do_step(0.05, 100)
    for i in (0, 5, T - 1):
        w = vals[i].copy()
        for _ in range(100):
            w = oracle.euler_do_step(w, 0.05, lambda q: ens.expr.eval(q, par[i]))
        assert got[i].tobytes() == w.tobytes()
    ens.close()


@pytest.mark.gpu
def test_program_with_transcendentals(gpu, oracle):
    """the pendulum, values [omega, theta, t], f = -(g / L) sin(theta): sin comes from the device's libm, so the comparison
    with the oracle's closure (host libm) is to 1e-13, not to the bit; EulerSolver.new mirrors EulerSolver::new(closure)"""
    from dzahui_b200 import euler as E
    T = 64
    theta0 = np.linspace(0.1, 3.0, T)
    vals = np.column_stack([np.zeros(T), theta0, np.zeros(T)])
    par = np.full((T, 1), 9.81 / 2.0)
    ens = E.EulerEnsemble.from_expression(lambda v, p: -p[0] * E.sin(v[1]), vals, par)
    got = ens.do_step(0.001, 1000)
    for i in (0, 31, T - 1):
        w = vals[i].copy()
        for _ in range(1000):
            w = oracle.euler_do_step(w, 0.001, lambda q: -par[i, 0] * math.sin(q[1]))
        assert np.allclose(got[i], w, rtol=1e-13, atol=1e-13)
    ens.close()
    s = E.EulerSolver.new(lambda val: -4.905 * E.sin(val[1]) + E.exp(-val[2]) * E.cos(val[2]) - E.log(2.0 + val[0] * val[0]), 3)
    w = np.array([0.0, 1.0, 0.0])
    state = list(w)
    for _ in range(50):
        state = s.do_step(state, 0.01)
        w = oracle.euler_do_step(w, 0.01, lambda q: -4.905 * math.sin(q[1]) + math.exp(-q[2]) * math.cos(q[2]) - math.log(2.0 + q[0] * q[0]))
    assert np.allclose(state, w, rtol=1e-13, atol=1e-13)


@pytest.mark.gpu
def test_malformed_programs_are_refused(gpu):
    import ctypes as C
    from dzahui_b200 import _lib, euler as E
    lib = _lib.load()
    vals = np.zeros((2, 2))

    def create(rpn, n_params=0):
        ops = (_lib.EulerOp * len(rpn))(*[_lib.EulerOp(E.OPCODES.get(op, op) if isinstance(op, str) else op, arg, val) for op, arg, val in rpn])
        par = np.zeros((2, max(n_params, 1)))
        h = C.c_void_p()
        rc = lib.dzb_euler_create_program(2, 2, C.cast(ops, C.c_void_p), len(rpn), par.ctypes.data_as(_lib.c_double_p), n_params,
                                          vals.ctypes.data_as(_lib.c_double_p), C.byref(h))
        if rc == 0:
            lib.dzb_euler_destroy(h)
        return rc

    assert create([("const", 0, 1.0)]) == 0
    assert create([("add", 0, 0.0)]) == 7                                   # pops an empty stack
    assert create([("const", 0, 1.0), ("const", 0, 2.0)]) == 7              # leaves two values
    assert create([("value", 2, 0.0)]) == 7                                 # [f64; 2] has no values[2]
    assert create([("param", 0, 0.0)]) == 7                                 # no parameters given
    assert create([(99, 0, 0.0)]) == 7                                      # unknown opcode
    assert create([("const", 0, 1.0)] * 17 + [("add", 0, 0.0)] * 16) == 7   # 17 stack entries
