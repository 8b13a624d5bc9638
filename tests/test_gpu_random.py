"""Seeded random sweeps of the whole path against the CPU oracle (GPU box): graded and jittered meshes, both signs of
the convection term, small and large quadrature orders (tabulated n <= 100 and asymptotic n >= 101 branches, down to a
single used node), ragged sizes.  FAITHFUL kernels must reproduce the oracle bit for bit; FAST / PARALLEL kernels must
stay within 1e-12 per band entry (plus a few ulps of the row scale where an entry's own terms cancel) and, for nodal values, within 1e-10 of the exact (__float128) solution of the
reference's system or within 4x the reference's own f64 error where that is larger (ill-conditioned draws)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def draw_mesh(rng, n):
    kind = rng.integers(0, 4)
    if kind == 0:
        x = np.linspace(0.0, rng.uniform(0.5, 20.0), n)
    elif kind == 1:
        x = np.cumsum(rng.uniform(0.5, 1.5, n))
    elif kind == 2:
        x = np.linspace(0.0, 1.0, n) ** rng.uniform(1.2, 3.0) * rng.uniform(1.0, 5.0)     # graded towards 0
    else:
        x = np.cumsum(np.exp(rng.normal(0.0, 0.7, n)))                                       # wildly uneven
    return x - x[0] + rng.uniform(-3.0, 3.0)


@pytest.fixture(scope="module")
def gpu(dz):
    assert dz.device_count() >= 1
    dz.set_device(0)
    return dz


@pytest.mark.parametrize("seed", range(12))
def test_random_static_diffusion(gpu, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        n = int(rng.choice([3, 4, 5, 9, 17, 33, 64, 100, 257, 1000, 2500]))
        G = int(rng.choice([2, 3, 7, 50, 100, 101, 150, 350]))
        mesh = draw_mesh(rng, n)
        mu, b = float(rng.uniform(0.2, 3.0)), float(rng.uniform(-2.0, 2.0))
        bc = (float(rng.normal()), float(rng.normal() * 5))
        lo, di, up, rhs = oracle.ti_assemble(mesh, mu, b, bc, G)
        ref = oracle.thomas(lo, di, up, rhs)
        p = gpu.DiffussionParams.time_independent().mu(mu).b(b).boundary_conditions(*bc).build()
        s = gpu.DiffussionSolverTimeIndependent(p, mesh, G, gpu.make_options(gpu.ASSEMBLY_FAITHFUL, gpu.SOLVE_FAITHFUL))
        glo, gdi, gup, grhs = s.bands()
        assert glo.tobytes() == lo.tobytes() and gdi.tobytes() == di.tobytes() and gup.tobytes() == up.tobytes(), (n, G)
        assert np.array_equal(s.solve(0.0), ref, equal_nan=True), (n, G)
        if not np.all(np.isfinite(ref)) or G < 3:
            continue   # a singular draw (e.g. one used node): the reference yields inf/NaN, and so did we, bit for bit
        truth = oracle.thomas(lo, di, up, rhs, kind="f128")
        e_ref = rel_l2(ref, truth)
        for amode in (gpu.ASSEMBLY_FAITHFUL, gpu.ASSEMBLY_FAST):
            s2 = gpu.DiffussionSolverTimeIndependent(p, mesh, G, gpu.make_options(amode, gpu.SOLVE_PARALLEL))
            b2 = s2.bands()
            scale = np.maximum.reduce([np.abs(lo), np.abs(di), np.abs(up)])
            # The reference evaluates a hat as c*(cs*x + ms) + t: the image cs*x + ms is rounded at the size of the
            # coordinate, then divided by the element length, so its own bands carry noise ~ eps*|x|/h per entry; the closed
            # form works with coordinate differences and cannot (and need not) reproduce that noise.
            h_loc = np.minimum(np.diff(mesh)[:-1], np.diff(mesh)[1:])
            noise = np.zeros(n)
            noise[1:-1] = 8 * np.finfo(float).eps * np.abs(mesh[1:-1]) / h_loc
            for k, (got, want) in enumerate(zip(b2[:3], (lo, di, up))):
                # 1e-12 of the entry (+ that noise), plus a few ulps of the row's largest entry where the entry's terms cancel
                ratio = np.abs(got - want) / ((1e-12 + noise) * np.abs(want) + (3e-14 + noise) * scale + 1e-300)
                i = int(np.argmax(ratio))
                assert ratio[i] <= 1.0, (n, G, amode, k, i, float(ratio[i]), float(got[i]), float(want[i]), float(scale[i]),
                                         mesh[max(i - 2, 0):i + 3].tolist(), mu, b)
            if amode == gpu.ASSEMBLY_FAITHFUL:   # same bands: the solve alone is under test
                e = rel_l2(s2.solve(0.0), truth)
                assert e <= max(1e-10, 4 * e_ref), (n, G, mu, b, e, e_ref)


@pytest.mark.parametrize("seed", range(8))
def test_random_time_dependent(gpu, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(4):
        n = int(rng.choice([3, 5, 12, 33, 64, 129, 300, 777, 1024]))
        G = int(rng.choice([7, 50, 100, 101, 150]))
        B = int(rng.integers(1, 9))
        meshes = np.stack([draw_mesh(rng, n) for _ in range(B)])
        mu, b = float(rng.uniform(0.2, 2.0)), float(rng.uniform(-1.5, 1.5))
        bc = (float(rng.normal()), float(rng.normal() * 3))
        ic = rng.normal(size=(B, n - 2))
        hmin = float(np.min(np.diff(meshes, axis=1)))
        dt = 0.05 * hmin * hmin / mu
        bands = oracle.ensemble_td_assemble(meshes, mu, b, G)
        st = np.stack([oracle.td_initial_state(ic[k], n, bc) for k in range(B)])
        want = oracle.ensemble_td_run(bands, st, dt, bc, 12)
        e = gpu.EnsembleTimeDependent(meshes, mu, b, bc, ic, G, gpu.make_options(gpu.ASSEMBLY_FAITHFUL, gpu.SOLVE_FAITHFUL))
        e.step(dt, 12)
        assert e.state().tobytes() == want.tobytes(), (n, G, B)
        for stored in (False, True):
            e2 = gpu.EnsembleTimeDependent(meshes, mu, b, bc, ic, G, gpu.make_options(gpu.ASSEMBLY_FAST, gpu.SOLVE_PARALLEL, -1, stored))
            for _ in range(4):
                e2.step(dt, 1)
            e2.step(dt, 8)
            got = e2.state()
            worst = max(rel_l2(got[k], want[k]) for k in range(B))
            assert worst <= 1e-10, (n, G, B, stored, worst)


@pytest.mark.parametrize("seed", range(4))
def test_random_stokes(gpu, oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    for _ in range(4):
        n = int(rng.choice([3, 4, 11, 30, 101]))
        G = int(rng.choice([7, 100, 101, 150, 350]))
        mesh = draw_mesh(rng, n)
        rho, pressure = float(rng.uniform(0.5, 3.0)), float(rng.normal() * 50)
        c = rng.normal(size=3)
        force = lambda x: c[0] + c[1] * x + c[2] * np.sin(x)
        lo, di, up, rhs = oracle.stokes_assemble(mesh, rho, pressure, G, force)
        ref = oracle.thomas(lo, di, up, rhs)
        p = gpu.StokesParams.normal_1d().hydrostatic_pressure(pressure).density(rho).force_function(force).build()
        s = gpu.StokesSolver1D(p, mesh, G)
        glo, gdi, gup, grhs = s.bands()
        assert glo.tobytes() == lo.tobytes() and gdi.tobytes() == di.tobytes() and gup.tobytes() == up.tobytes()
        assert grhs.tobytes() == rhs.tobytes()
        assert np.array_equal(s.solve(0.0), ref, equal_nan=True)


@pytest.mark.parametrize("scale", [1e-75, 1e-30, 1e-3, 1.0, 1e3, 1e30, 1e75])
@pytest.mark.parametrize("bc0", [0.0, -0.0, 1.0])
def test_reference_arithmetic_bits_across_magnitudes(gpu, oracle, scale, bc0):
    """The warp-per-bar sweeps divide by a stored correctly rounded reciprocal (exact_div) inside a safe range and fall back
    to the division outside it: bars whose element lengths span 150 decades put the pivots (~1/h for the stiffness, ~h
    for the mass matrix) on both sides of that range, zero left boundary values make every forward numerator a signed zero.
    Static and time-dependent FAITHFUL results must keep the oracle's bits in all of them."""
    rng = np.random.default_rng(int(abs(np.log10(scale)) * 7) + (3 if bc0 else 0))
    n = 777
    mesh = np.cumsum(rng.uniform(0.5, 1.5, n)) * scale
    opt = gpu.make_options(gpu.ASSEMBLY_FAITHFUL, gpu.SOLVE_FAITHFUL)
    mu, b = 1.0, 0.25 / scale  # keeps the cell Peclet number b h / (2 mu) of order 0.1 at every scale
    p = gpu.DiffussionParams.time_independent().mu(mu).b(b).boundary_conditions(bc0, 15.0).build()
    s = gpu.DiffussionSolverTimeIndependent(p, mesh, 40, opt)
    want = oracle.thomas(*oracle.ti_assemble(mesh, mu, b, (bc0, 15.0), 40))
    got = s.solve(0.0)
    assert got.tobytes() == want.tobytes()
    ic = rng.uniform(-1.0, 1.0, n - 2)
    q = gpu.DiffussionParams.time_dependent().mu(mu).b(b).boundary_conditions(bc0, 15.0).initial_conditions(ic).build()
    td = gpu.DiffussionSolverTimeDependent(q, mesh, 40, opt)
    dt = 1e-3 * scale * scale
    for _ in range(5):
        u = td.solve(dt)
    wantu = oracle.td_run(oracle.td_assemble(mesh, mu, b, 40), oracle.td_initial_state(ic, n, (bc0, 15.0)), dt, (bc0, 15.0), 5)
    assert u.tobytes() == wantu.tobytes()
