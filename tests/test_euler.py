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
