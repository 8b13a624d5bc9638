"""CPU-only: pins the oracle (oracle/, the C++ restatement of the reference) against
 (a) the reference's own 17 hot-path unit tests, restated with their exact asserted windows, and
 (b) the restatement-derived golden vectors of SURVEY.md §8(c) and tests/golden/golden.json.
The reference (Rust) cannot be built in this image, so its tests' windows are the only reference-held pins."""
import math

import numpy as np
import pytest

import oracle_lib as o


# ---------------------------------------------------------------- quadrature/gauss_legendre.rs tests :5938-6003
def _integrate(n, f):
    return sum(w * f(math.cos(th)) for th, w in (o.quad_pair(n, k) for k in range(1, n)))


def test_integrate_exp_tabulate():                    # :5938  n = 100, tabulated
    assert abs(_integrate(100, math.exp) - (math.e - 1 / math.e)) <= 1e-3


def test_integrate_exp_without_tabulation():          # :5951  n = 600, calculated
    assert abs(_integrate(600, math.exp) - (math.e - 1 / math.e)) <= 1e-5


def test_integrate_cosine_tabulate():                 # :5964  n = 50
    assert abs(_integrate(50, math.cos) - 2 * math.sin(1.0)) <= 1e-2


def test_integrate_cosine_without_tabulation():       # :5977  n = 510
    assert abs(_integrate(510, math.cos) - 2 * math.sin(1.0)) <= 1e-4


def test_integrate_cosine_0_pi():                     # :5990  affine map to [0, pi], n = 510
    cs, ms = (math.pi - 0.0) / 2.0, (math.pi + 0.0) / 2.0
    s = sum(w * math.cos(cs * math.cos(th) + ms) * cs for th, w in (o.quad_pair(510, k) for k in range(1, 510)))
    assert abs(s) <= 1e-4


def test_quad_pair_rejects_k_equal_n():               # :5908-5926
    for n in (5, 100, 101, 350):
        with pytest.raises(o.OracleError) as e:
            o.quad_pair(n, n)
        assert e.value.code == o.INTEGRATION
        o.quad_pair(n, n - 1)


def test_quadrature_sanity_values():
    for G, sw, x1, xl in [(50, 1.9970913774468448, 0.998866404420071, -0.9940319694320907),
                          (150, 1.9996723913294463, 0.9998723404457334, -0.9993274305065947),
                          (350, 1.999939595579857, 0.9999764625657942, -0.9998759847381288)]:
        x, w = o.quad_table(G)
        assert abs(math.fsum(w) - sw) < 1e-14 and x[0] == x1 and x[-1] == xl
        assert np.all(np.diff(x) < 0)                 # strictly decreasing in k
    assert abs(math.fsum(o.quad_table(100)[1]) - 1.9992653655094943) < 1e-14
    assert abs(math.fsum(o.quad_table(600)[1]) - 1.9999794211462103) < 1e-14


# ---------------------------------------------------------------- linear_basis.rs:103 transform_basis_three_nodes
def test_transform_basis_three_nodes():
    b = o.linear_basis([0.0, 1.0, 2.0])
    assert len(b) == 3
    assert b[0] == ([0.0, -1.0, 0.0], [0.0, 1.0, 0.0], [0.0, 1.0])
    assert b[1] == ([0.0, 1.0, -1.0, 0.0], [0.0, 0.0, 2.0, 0.0], [0.0, 1.0, 2.0])
    assert b[2] == ([0.0, 1.0, 0.0], [0.0, -1.0, 0.0], [1.0, 2.0])


def test_piecewise_strict_less_than_selection():      # piecewise_polynomials_1degree.rs:133-148
    mesh = [0.0, 1.0, 2.0]
    assert o.basis_evaluate(mesh, 1, 1.0) == 1.0      # x == breakpoint -> next piece (right): -1*1 + 2
    assert o.basis_evaluate(mesh, 1, 2.0) == 0.0      # beyond the last breakpoint -> last (zero) piece
    assert o.basis_evaluate(mesh, 1, -0.5) == 0.0
    assert o.basis_evaluate(mesh, 1, 0.25, derivative=True) == 1.0
    assert o.basis_evaluate(mesh, 1, 1.25, derivative=True) == -1.0


# ---------------------------------------------------------------- matrix_solver/mod.rs:58,72
def test_solve_3x3():
    r = o.thomas_dense([[1., 2., 0.], [1., 1., 2.], [0., 2., 1.]], [1., 0., 0.])
    assert 0.5 <= r[0] <= 0.7 and 0.1 <= r[1] <= 0.3 and -0.5 <= r[2] <= -0.3


def test_solve_5x5():
    a = [[1., 2., 0., 0., 0.], [2., 1., 1., 0., 0.], [0., 1., 2., 1., 0.], [0., 0., 2., 2., 1.], [0., 0., 0., 1., 2.]]
    r = o.thomas_dense(a, [1., 0., 0., 0., 0.])
    assert 0.09 <= r[0] <= 0.13 and 0.43 <= r[1] <= 0.46 and -0.70 <= r[2] <= -0.60
    assert 0.86 <= r[3] <= 0.90 and -0.46 <= r[4] <= -0.42


def test_thomas_wrong_dims():                         # mod.rs:19-21
    with pytest.raises(o.OracleError) as e:
        o.thomas_dense(np.zeros((3, 2)), [1., 0., 0.])
    assert e.value.code == o.WRONG_DIMS
    with pytest.raises(o.OracleError):
        o.thomas_dense(np.eye(3), [1., 0.])
    with pytest.raises(o.OracleError):                # utils.rs:43-45 and :18-20
        o.tridiag_matvec([0, 1, 1], [1, 1, 1], [1, 1, 0], [1, 2], 1.0)
    with pytest.raises(o.OracleError):
        o.add([1., 2.], [1.])


# ---------------------------------------------------------------- time_independent.rs tests :207-335
def _ti(mesh, mode=0):
    lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (0.0, 1.0), 150, mode=mode)
    return lo, di, up, rhs, o.thomas(lo, di, up, rhs)


def test_regular_mesh_matrix_3p():
    lo, di, up, rhs, _ = _ti([0.0, 0.5, 1.0])
    assert di[0] == 1.0 and di[2] == 1.0
    assert 3.9 <= di[1] <= 4.1 and up[1] <= -1.4 and lo[1] <= -2.4
    assert abs(di[1] - 3.99934) < 1e-5 and abs(up[1] + 1.49984) < 1e-5 and abs(lo[1] + 2.49967) < 1e-5


def test_solve_system_3p():
    r = _ti([0.0, 0.5, 1.0])[4]
    assert len(r) == 3 and 0.2 <= r[1] <= 0.4 and abs(r[1] - 0.375020) < 1e-6


def test_regular_mesh_matrix_4p():
    lo, di, up, rhs, _ = _ti([0.0, 0.33, 0.66, 1.0])
    assert lo[1] <= -3.4 and up[1] >= -3.6 and 5.9 <= di[1] <= 6.1 and -2.6 <= up[1] <= -2.4
    assert -3.6 <= lo[2] <= -3.4 and 5.9 <= di[2] <= 6.1 and up[2] <= -2.3 and di[2] >= -2.5
    assert di[0] == 1.0 and di[3] == 1.0


def test_solve_system_4p():
    r = _ti([0.0, 0.33, 0.66, 1.0])[4]
    assert len(r) == 4 and r[0] == 0.0 and 0.20 <= r[1] <= 0.24 and 0.52 <= r[2] <= 0.56 and r[3] == 1.0
    assert abs(r[1] - 0.227284) < 1e-6 and abs(r[2] - 0.544376) < 1e-6


def test_regular_mesh_bigger_matrix():
    lo, di, up, rhs, _ = _ti([0.0, 0.25, 0.5, 0.75, 1.0])
    for i in (1, 2, 3):
        assert -4.6 <= lo[i] <= -4.4 and 7.9 <= di[i] <= 8.1 and -3.6 <= up[i] <= -3.4
    assert di[0] == 1.0 and di[4] == 1.0 and rhs[0] == 0.0 and rhs[4] == 1.0


def test_solve_bigger_system():
    r = _ti([0.0, 0.25, 0.5, 0.75, 1.0])[4]
    assert len(r) == 5 and r[0] == 0.0 and r[4] == 1.0
    assert 0.15 <= r[1] <= 0.17 and 0.36 <= r[2] <= 0.38 and 0.63 <= r[3] <= 0.655
    assert abs(r[1] - 0.164922) < 1e-6 and abs(r[2] - 0.376956) < 1e-6 and abs(r[3] - 0.649552) < 1e-6


# ---------------------------------------------------------------- time_dependent.rs tests :290,:320
def test_matrix_and_vector_values_3p():
    mlo, mdi, mup, slo, sdi, sup = o.td_assemble([0.0, 0.5, 1.0], 1.0, 1.0, 150)
    assert mdi[0] == 1.0 and mdi[2] == 1.0 and sdi[0] == 1.0 and sdi[2] == 1.0
    assert 0.08 <= mlo[1] <= 0.09 and 0.3 <= mdi[1] <= 0.35 and 0.08 <= mup[1] <= 0.09
    assert 2.4 <= slo[1] <= 2.6 and -4.1 <= sdi[1] <= -3.9 and 1.4 <= sup[1] <= 1.6


def test_matrix_solved_3p():
    bands = o.td_assemble([0.0, 0.5, 1.0], 1.0, 1.0, 150)
    st = o.td_run(bands, o.td_initial_state([15.0], 3, (1.0, 0.0)), 0.01, (1.0, 0.0), 1000)
    assert st[0] == 1.0 and st[2] == 0.0 and 0.55 <= st[1] <= 0.65
    assert abs(st[1] - 0.6041837374245905) <= 1e-13


def test_td_wrong_dims():                             # time_dependent.rs:66-68
    with pytest.raises(o.OracleError) as e:
        o.td_initial_state([1.0, 2.0], 3, (0.0, 1.0))
    assert e.value.code == o.WRONG_DIMS


# ---------------------------------------------------------------- stokes_solver/dim1.rs:264 regular_mesh_matrix_4p_nav
def test_regular_mesh_matrix_4p_nav():
    lo, di, up, rhs = o.stokes_assemble([0.0, 0.333, 0.666, 1.0], 1.0, 1.0, 150, lambda _x: 10.0)
    assert -0.6 <= di[0] <= -0.4 and 0.4 <= up[0] <= 0.6
    assert -0.6 <= lo[1] <= -0.4 and -0.1 <= di[1] <= 0.1 and 0.4 <= up[1] <= 0.6
    assert -0.6 <= lo[2] <= -0.4 and -0.1 <= di[2] <= 0.1 and 0.4 <= up[2] <= 0.6
    assert lo[3] == 0.0 and di[3] == 1.0
    assert 1.55 <= rhs[0] <= 1.75 and 3.25 <= rhs[1] <= 3.45 and 3.25 <= rhs[2] <= 3.45 and rhs[3] == 1.0
    sol = o.thomas(lo, di, up, rhs)
    assert -9.1 <= sol[0] <= -8.9 and -5.7 <= sol[1] <= -5.5 and -2.4 <= sol[2] <= -2.2


# ---------------------------------------------------------------- SURVEY §8(c) goldens (must hold to <= 1e-13 relative)
def test_survey_golden_c1(golden, assets_dir):
    m = o.mesh_ingest_1d(assets_dir + "/1dbar.obj")
    mesh = o.filter_for_solving_1d(m["vertices"])
    lo, di, up, rhs = o.stokes_assemble(mesh, 1.0, 100.0, 350, -10.0)
    p = o.thomas(lo, di, up, rhs)
    gold = [109.96852572142232, 108.96852572142232, 107.97506144001291, 106.9750010305406, 105.98147673444058,
            104.98135591184699, 103.98777159743146, 102.98759035806665, 101.9939460217097, 100.9937043619243, 100.0]
    assert np.max(np.abs(p - gold) / np.abs(gold)) <= 1e-13
    assert abs(di[0] + 0.4999697981453693) <= 1e-15 and abs(up[0] - 0.4999697981453693) <= 1e-15
    assert abs(rhs[0] + 0.4999697981453693) <= 1e-15 and abs(lo[1] + 0.49999999964455916) <= 1e-15
    assert abs(up[1] - 0.4999697981453691) <= 1e-15 and abs(rhs[1] + 0.9999933036543858) <= 1e-15
    assert np.array_equal(p, golden["C1"]["solution"])


def test_survey_golden_c2(golden, assets_dir):
    m = o.mesh_ingest_1d(assets_dir + "/1dbar_many_divisions_irregular.obj", 50.0)
    mesh = o.filter_for_solving_1d(m["vertices"])
    assert len(mesh) == 78 and mesh[-1] == 8.067
    lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150)
    u = o.thomas(lo, di, up, rhs)
    head = [1.0, 0.9485126528155906, 0.9131288832361018, 0.8732300262353963, 0.8111952905268778]
    tail = [8.310922166290624, 9.523551436534156, 11.470321194454337, 12.838339700671387, 15.0]
    assert np.max(np.abs(u[:5] - head) / np.abs(head)) <= 1e-13 and np.max(np.abs(u[-5:] - tail) / np.abs(tail)) <= 1e-13
    assert abs(di[1] - 114.62950072523479) <= 1e-12
    assert np.array_equal(u, golden["C2"]["solution"])


def test_golden_c3(golden):
    g = golden["C3"]
    mesh = np.array(g["mesh"])
    assert len(mesh) == 11 and mesh[0] == -35.710892 and mesh[5] == 0.0 and mesh[-1] == 35.710892
    bands = o.td_assemble(mesh, 1.0, 1.0, 150)
    st = o.td_run(bands, o.td_initial_state(np.zeros(9), 11, (1.0, 15.0)), 1e-3, (1.0, 15.0), 10000)
    assert np.array_equal(st, g["states"]["10000"])
    approx = [-0.771, -0.255, -0.273, 0.553, -1.291, 3.006, -6.793, 14.823, -31.000]
    assert np.max(np.abs(st[1:-1] - approx)) < 1e-3


# ---------------------------------------------------------------- storage / recomputation modes agree bit for bit
@pytest.mark.parametrize("mode", [1, 2])
def test_dense_and_recompute_modes_bit_identical(mode, golden):
    mesh = np.array(golden["C2"]["mesh"])
    a = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150, mode=0)
    b = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150, mode=mode)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    a = o.td_assemble(mesh[:20], 1.0, 1.0, 150, mode=0)
    b = o.td_assemble(mesh[:20], 1.0, 1.0, 150, mode=mode)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    a = o.stokes_assemble(mesh[:12], 1.0, 3.0, 101, lambda x: x * x, mode=0)
    b = o.stokes_assemble(mesh[:12], 1.0, 3.0, 101, lambda x: x * x, mode=mode)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_dense_td_run_matches_banded(golden):
    g = golden["unit"]["td_3p"]
    st = o.td_run(g["bands"], o.td_initial_state([15.0], 3, (1.0, 0.0)), 0.01, (1.0, 0.0), 1000, dense=True)
    assert np.array_equal(st, g["state_1000"])


def test_openmp_assembly_bit_identical(golden):
    mesh = np.array(golden["C2"]["mesh"])
    a = o.td_assemble(mesh, 1.0, 1.0, 150, threads=1)
    b = o.td_assemble(mesh, 1.0, 1.0, 150, threads=4)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_extended_precision_thomas_close(golden):
    g = golden["C2"]
    u = o.thomas(g["lo"], g["di"], g["up"], g["rhs"])
    for kind in ("f80", "f128"):
        t = o.thomas(g["lo"], g["di"], g["up"], g["rhs"], kind=kind)
        assert np.linalg.norm(u - t) / np.linalg.norm(t) < 1e-13


def test_exact_arithmetic_system_closed_form_equals_node_loop():
    """dzo_ti_intended_f128 (the reference's static system, time_independent.rs:91-182, in __float128 on its f64 inputs): the
    moment form the long-bar yardstick uses is the node-by-node loop, and the f64 oracle's bands are its bands up to f64
    rounding of a 149- / 349-term sum."""
    rng = np.random.default_rng(3)
    for n, G in [(50, 150), (200, 150), (37, 350), (30, 20)]:
        x = np.sort(rng.uniform(0.0, 3.0, n))
        ub, lb, db, pb = o.ti_intended_f128(x, 1.3, 0.7, (1.0, 15.0), G, brute=True, bands=True)
        uc, lc, dc, pc = o.ti_intended_f128(x, 1.3, 0.7, (1.0, 15.0), G, brute=False, bands=True, threads=2)
        assert np.array_equal(ub, uc) and np.array_equal(lb, lc) and np.array_equal(db, dc) and np.array_equal(pb, pc)
        lo, di, up, rhs = o.ti_assemble(x, 1.3, 0.7, (1.0, 15.0), G)
        for got, want in ((lo, lc), (di, dc), (up, pc)):
            assert np.max(np.abs(got[1:-1] - want[1:-1]) / np.abs(want[1:-1])) < 5e-15
        assert dc[0] == 1.0 and dc[-1] == 1.0 and lc[0] == 0.0 and pc[-1] == 0.0
        u64 = o.thomas(lo, di, up, rhs)
        assert np.linalg.norm(u64 - uc) / np.linalg.norm(uc) < 1e-11  # short bars: the reference is at its intended system


def test_exact_arithmetic_mass_and_stiffness_agree_with_the_oracle_loop():
    """dzo_td_intended_f128: M and S of the time-dependent solver (time_dependent.rs:110-247) from the rule's moments in
    __float128 — an independent derivation of what the line-by-line restatement sums node by node.  Moment form == node loop
    (identical after rounding); the f64 oracle agrees to the rounding of a (G-1)-term sum, per row relative to the row's scale."""
    rng = np.random.default_rng(4)
    for n, G in [(60, 150), (33, 350), (25, 20), (40, 101)]:
        x = np.sort(rng.uniform(-2.0, 3.0, n))
        brute = o.td_intended_f128(x, 1.3, 0.7, G, brute=True)
        exact = o.td_intended_f128(x, 1.3, 0.7, G, brute=False, threads=2)
        assert all(np.array_equal(p, q) for p, q in zip(brute, exact))
        got = o.td_assemble(x, 1.3, 0.7, G)
        for k in (0, 3):  # M bands, S bands
            scale = sum(np.abs(exact[k + j]) for j in range(3))
            for j in range(3):
                # (phi is evaluated as slope * x + intercept: an absolute error of eps |x| / h per node, |x| / h up to ~50 here)
                assert np.max(np.abs(got[k + j] - exact[k + j]) / scale) < 2e-13
        for band, end in zip(exact, (0.0, 1.0, 0.0, 0.0, 1.0, 0.0)):
            assert band[0] == end and band[-1] == end      # identity rows at both ends (:236-239)


def test_reference_f64_result_is_noise_limited_on_a_long_bar():
    """On BASELINE config 4's regular bar the reference's own f64 result is far from the system it means to solve — and solving
    its f64 bands exactly does not help: the distance is the rounding noise of the band entries."""
    n = 200_000
    x = np.linspace(0.0, 1.0, n)
    truth = o.ti_intended_f128(x, 1.0, 1.0, (1.0, 15.0), 150, threads=o.hardware_threads())
    lo, di, up, rhs = o.ti_assemble(x, 1.0, 1.0, (1.0, 15.0), 150, threads=o.hardware_threads())
    e64 = np.linalg.norm(o.thomas(lo, di, up, rhs) - truth) / np.linalg.norm(truth)
    e128 = np.linalg.norm(o.thomas(lo, di, up, rhs, kind="f128") - truth) / np.linalg.norm(truth)
    assert 1e-8 < e64 < 1e-3 and abs(e64 - e128) < 0.05 * e64


# ---------------------------------------------------------------- ingestion (mesh_builder.rs:204-328, mesh/mod.rs:54-68)
def test_ingest_1dbar(golden, assets_dir):
    m = o.mesh_ingest_1d(assets_dir + "/1dbar.obj")
    n = m["n"]
    assert n == 11 and len(m["vertices"]) == 132 and len(m["indices"]) == 60
    v = m["vertices"].reshape(2 * n, 6)
    assert np.array_equal(v[:n, 0], [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])   # kept axis: z
    assert np.all(v[:n, 1] == 0.0) and np.all(v[n:, 1] == m["prom_width"]) and np.all(v[:, 5] == 1.0)
    assert m["max_length"] == 1.0 and m["prom_width"] == 1.0 * 6.0 / (66.0 - 6.0)
    assert list(m["indices"][:6]) == [0, 1, 11, 10, 21, 20]
    assert list(m["indices"][6:12]) == [1, 12, 11, 1, 2, 12]
    assert np.array_equal(m["vertices"], golden["ingest"]["1dbar.obj"]["vertices"])


def test_ingest_rejects_planar_grid(assets_dir):      # big_mesh.obj: x and z both vary (mesh_builder.rs:217-227)
    with pytest.raises(o.OracleError) as e:
        o.mesh_ingest_1d(assets_dir + "/big_mesh.obj")
    assert e.value.code == o.MESH_PARSE


def test_ingest_errors_and_ordering(tmp_path):
    p = tmp_path / "m.obj"
    with pytest.raises(o.OracleError) as e:
        o.mesh_ingest_1d(str(tmp_path / "missing.obj"))
    assert e.value.code == o.IO
    p.write_text("v 0.0 0.0 1.0\nv 0.0 0.0\n")
    with pytest.raises(o.OracleError):
        o.mesh_ingest_1d(str(p))
    p.write_text("v 0.0 0.0 abc\nv 0.0 0.0 1.0\n")
    with pytest.raises(o.OracleError):
        o.mesh_ingest_1d(str(p))
    p.write_text("v 0.0  0.0 1.0\n")                  # double blank -> empty token -> parse error
    with pytest.raises(o.OracleError):
        o.mesh_ingest_1d(str(p))
    # unsorted input, y varies, -0.0 / 0.0 keep file order, CRLF tolerated, non-vertex lines ignored
    p.write_text("# c\nv 1.5 3.0 2.0\r\nvn 0 0 1\nv 1.5 -1.0 2.0\nv 1.5 0.0 2.0\nv 1.5 -0.0 2.0\nv 1.5 1e0 2.0\nf 1 2 3\n")
    m = o.mesh_ingest_1d(str(p))
    nodes = o.filter_for_solving_1d(m["vertices"])
    assert list(nodes) == [-1.0, 0.0, 0.0, 1.0, 3.0]
    assert not np.signbit(nodes[1]) and np.signbit(nodes[2])


def test_quadrature_tables_equal_the_reference_literals(tmp_path):
    """Build container only (the reference tree is not on the GPU box): the 201 Gauss-Legendre tables the oracle and the
    product are compiled with (regenerated from first principles by tools/gen_gl_tables.py) equal the reference's 25-30
    digit literals (quadrature/gauss_legendre.rs:22-5673) bit for bit as f64, and both committed copies are that output."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir("/root/reference/src/lib/solvers/quadrature"):
        pytest.skip("reference tree not present")
    try:
        import mpmath  # noqa: F401
    except ImportError:
        pytest.skip("mpmath not installed")
    out = tmp_path / "gl.h"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_gl_tables.py"), "--check", "/root/reference", "--out", str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 mismatches" in r.stdout
    gen = out.read_bytes()
    assert gen == open(os.path.join(root, "oracle", "gl_tables_generated.h"), "rb").read()
    assert gen == open(os.path.join(root, "dzahui_b200", "csrc", "gl_tables_generated.h"), "rb").read()


# ---------------------------------------------------------------- 2D ingestion, mesh_builder.rs:342-500
def test_oracle_2d_ingestion_known_answers(assets_dir, golden, tmp_path):
    """independent pins of the restated build_mesh_2d: every vertex of a lone triangle / a two-triangle trapezoid is on
    the boundary; the boundary of big_mesh.obj (an 11 x 11 planar grid, 200 triangles) is its 40 perimeter vertices"""
    t = o.mesh_ingest_2d(f"{assets_dir}/test.obj")
    assert list(t["boundary_indices"]) == [0, 1, 2] and list(t["indices"]) == [0, 1, 2]
    assert list(t["vertices"]) == [-1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    assert t["max_length"] == 2.0 and list(t["translation"]) == [0.0, 1.0, 0.0]     # x_min + 2/2, y_min(=0.0) + 2/2
    z = o.mesh_ingest_2d(f"{assets_dir}/trapezoid.obj")
    assert list(z["boundary_indices"]) == [0, 1, 2, 3]
    b = o.mesh_ingest_2d(f"{assets_dir}/big_mesh.obj")
    assert (b["n_vertices"], b["n_triangles"]) == (121, 200)
    xy = b["vertices"].reshape(-1, 6)[:, :2]
    on_edge = (xy[:, 0] == xy[:, 0].min()) | (xy[:, 0] == xy[:, 0].max()) | (xy[:, 1] == xy[:, 1].min()) | (xy[:, 1] == xy[:, 1].max())
    assert list(b["boundary_indices"]) == list(np.nonzero(on_edge)[0]) and len(b["boundary_indices"]) == 40
    assert np.all(b["vertices"].reshape(-1, 6)[:, 2:] == [0.0, 0.0, 0.0, 1.0])
    for name, g in golden["ingest2d"].items():
        m = o.mesh_ingest_2d(f"{assets_dir}/{name}")
        assert list(m["boundary_indices"]) == g["boundary_indices"] and list(m["indices"]) == g["indices"]
        assert m["vertices"].tobytes() == np.array(g["vertices"]).tobytes() and m["max_length"] == g["max_length"]
    # the in-memory variant agrees with the file variant
    assert list(o.boundary_vertices(b["indices"].reshape(-1, 3))) == list(b["boundary_indices"])
    nofaces = tmp_path / "nofaces.obj"
    nofaces.write_text("v 0.5 0.0 1.0\nv 1.0 0.5 1.0\nv 0.25 1.0 1.0\n")
    with pytest.raises(o.OracleError):          # no `f` lines: no boundary edge (the reference panics inside merge_sort)
        o.mesh_ingest_2d(str(nofaces))
