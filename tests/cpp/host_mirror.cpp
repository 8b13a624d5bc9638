// host_mirror.cpp — the reference's own hot-path unit tests, re-stated against the C++ host mirror (dzahui_b200.hpp).
//   host_mirror --cpu               builders / panics / error vocabulary / quad_pair / parser errors / Writer (no GPU needed)
//   host_mirror --gpu <assets_dir>  the solver tests of the reference with their asserted windows, on the device
// Each block names the Rust test it restates (paths under /root/reference/src/lib).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "dzahui_b200.hpp"

using namespace dzahui;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      ++failures;                                                          \
    }                                                                      \
  } while (0)

template <class F>
static bool panics_with(F&& f, const char* msg) {
  try {
    f();
  } catch (const Panic& p) {
    return std::string(p.what()) == msg;
  }
  return false;
}

static void cpu_tests(const std::string& tmp) {
  // builders: stokes_solver/mod.rs:70-119, diffusion_solver/mod.rs:60-181
  auto p = StokesParams::static_pressure().hydrostatic_pressure(100.0).density(1.0).force_function([](double) { return -10.0; }).build();
  CHECK(p.rho == 1.0 && p.hydrostatic_pressure == 100.0 && p.force_function(3.0) == -10.0);
  CHECK(panics_with([] { StokesParams::normal_1d().density(1.0).force_function([](double x) { return x; }).build(); }, "Params lack 'pressure' term!"));
  CHECK(panics_with([] { StokesParams::normal_1d().hydrostatic_pressure(1.0).force_function([](double x) { return x; }).build(); }, "Params lack 'density' term!"));
  CHECK(panics_with([] { StokesParams::normal_1d().hydrostatic_pressure(1.0).density(1.0).build(); }, "Params lack force_function!"));
  auto q = DiffussionParams::time_independent().b(1.0).mu(2.0).boundary_conditions(0.0, 1.0).build();
  CHECK(q.mu == 2.0 && q.b == 1.0 && q.boundary_conditions[1] == 1.0);
  CHECK(panics_with([] { DiffussionParams::time_independent().b(1.0).boundary_conditions(0, 1).build(); }, "Params lack 'mu' term!"));
  CHECK(panics_with([] { DiffussionParams::time_independent().mu(1.0).boundary_conditions(0, 1).build(); }, "Params lack 'b' term!"));
  CHECK(panics_with([] { DiffussionParams::time_independent().mu(1.0).b(1.0).build(); }, "Params lack boundary conditions!"));
  CHECK(panics_with([] { DiffussionParams::time_dependent().mu(1.0).b(1.0).boundary_conditions(1, 0).build(); }, "Params lack initial conditions!"));
  // quad_pair misuse: quadrature/gauss_legendre.rs:5923-5925
  auto bad = quad_pair(150, 150);
  CHECK(bad.is_err() && bad.error().kind == Error::Integration);
  CHECK(bad.error().to_string() == "Error on integration method occurred: Misuse of quad_pair function: k should be smaller than n");
  CHECK(quad_pair(150, 149).is_ok());
  {  // integrate_cosine_tabulate (gauss_legendre.rs:5964): n = 50, err <= 1e-2
    double s = 0;
    for (size_t k = 1; k < 50; ++k) {
      auto pw = quad_pair(50, k).unwrap();
      s += pw.second * std::cos(std::cos(pw.first));
    }
    CHECK(std::fabs(s - 2 * std::sin(1.0)) <= 1e-2);
  }
  // parser errors are host code: mesh_builder.rs:58-84, :217-227
  auto missing = Mesh::builder(tmp + "/nope.obj").build_mesh_1d();
  CHECK(missing.is_err() && missing.error().kind == Error::Io);
  {
    std::ofstream f(tmp + "/bad.obj");
    f << "v 0.0 0.0 1.0\nv 0.0 0.0\n";
  }
  auto parse = Mesh::builder(tmp + "/bad.obj").build_mesh_1d();
  CHECK(parse.is_err() && parse.error().kind == Error::MeshParse);
  CHECK(parse.error().to_string().rfind("Unable to parse mesh file: ", 0) == 0);
  // Writer: writer.rs:47-146
  auto w = Writer::create(tmp, "p_", {"p"}, false).unwrap();
  CHECK(w.write(2.5, {109.96852572142232, 100.0, 1e-7}).is_ok());
  std::ifstream in(tmp + "/p_2.5.csv");
  std::stringstream ss;
  ss << in.rdbuf();
  CHECK(ss.str() == "p\n109.96852572142232\n100\n0.0000001\n");
  CHECK(Writer::create(tmp + "/missing", "a", {"p"}, false).is_err());
  // no CPU fallback: without a device every compute entry point reports an error instead of computing on the host
  int devices = 0;
  dzb_device_count(&devices);
  if (devices == 0) CHECK(Mesh::from_nodes({0.0, 0.5, 1.0}).is_err());
}

static bool within(double v, double lo, double hi) { return v >= lo && v <= hi; }

static void gpu_tests(const std::string& assets) {
  {  // time_independent.rs:207 regular_mesh_matrix_3p and :226 solve_system_3p
    auto p = DiffussionParams::time_independent().mu(1.0).b(1.0).boundary_conditions(0.0, 1.0).build();
    auto s = DiffussionSolverTimeIndependent::create(p, std::vector<double>{0.0, 0.5, 1.0}, 150).unwrap();
    auto K = s.stiffness_matrix();
    CHECK(within(K.at(1, 1), 3.9, 4.1) && K.at(1, 2) <= -1.4 && K.at(1, 0) <= -2.4);
    auto u = s.solve(0.0).unwrap();
    CHECK(within(u[1], 0.2, 0.4) && u[0] == 0.0 && u[2] == 1.0);
    CHECK(std::fabs(u[1] - 0.375020) < 1e-6);  // SURVEY §4 restated value
  }
  {  // :267 solve_system_4p
    auto p = DiffussionParams::time_independent().mu(1.0).b(1.0).boundary_conditions(0.0, 1.0).build();
    auto u = DiffussionSolverTimeIndependent::create(p, std::vector<double>{0.0, 0.33, 0.66, 1.0}, 150).unwrap().solve(0.0).unwrap();
    CHECK(within(u[1], 0.20, 0.24) && within(u[2], 0.52, 0.56));
  }
  {  // :316 solve_bigger_system
    auto p = DiffussionParams::time_independent().mu(1.0).b(1.0).boundary_conditions(0.0, 1.0).build();
    auto u = DiffussionSolverTimeIndependent::create(p, std::vector<double>{0.0, 0.25, 0.5, 0.75, 1.0}, 150).unwrap().solve(0.0).unwrap();
    CHECK(within(u[1], 0.15, 0.17) && within(u[2], 0.36, 0.38) && within(u[3], 0.63, 0.655));
  }
  {  // time_dependent.rs:290 test_matrix_and_vector_values_3p and :320 test_matrix_solved_3p
    auto p = DiffussionParams::time_dependent().mu(1.0).b(1.0).boundary_conditions(1.0, 0.0).initial_conditions({15.0}).build();
    auto s = DiffussionSolverTimeDependent::create(p, std::vector<double>{0.0, 0.5, 1.0}, 150).unwrap();
    auto M = s.mass_matrix(), S = s.stiffness_matrix();
    CHECK(within(M.at(1, 0), 0.08, 0.09) && within(M.at(1, 1), 0.30, 0.35));
    CHECK(within(S.at(1, 0), 2.4, 2.6) && within(S.at(1, 1), -4.1, -3.9) && within(S.at(1, 2), 1.4, 1.6));
    std::vector<double> u;
    for (int k = 0; k < 1000; ++k) u = s.solve(0.01).unwrap();
    CHECK(u[0] == 1.0 && u[2] == 0.0 && within(u[1], 0.55, 0.65));
    CHECK(std::fabs(u[1] - 0.6041837374245905) < 1e-12);
    // WrongDims: time_dependent.rs:66-68
    auto bad = DiffussionParams::time_dependent().mu(1.0).b(1.0).boundary_conditions(1.0, 0.0).initial_conditions({1.0, 2.0}).build();
    auto r = DiffussionSolverTimeDependent::create(bad, std::vector<double>{0.0, 0.5, 1.0}, 150);
    CHECK(r.is_err() && r.error().kind == Error::WrongDims);
    CHECK(r.error().to_string() == "One or more of the provided elements do not have the correct dimensions");
  }
  {  // dim1.rs:264 regular_mesh_matrix_4p_nav
    auto p = StokesParams::normal_1d().hydrostatic_pressure(1.0).density(1.0).force_function([](double) { return 10.0; }).build();
    auto s = StokesSolver1D::create(p, std::vector<double>{0.0, 0.333, 0.666, 1.0}, 150).unwrap();
    auto K = s.stiffness_matrix();
    auto b = s.b_vector();
    CHECK(within(K.at(0, 0), -0.6, -0.4) && within(K.at(0, 1), 0.4, 0.6) && within(K.at(1, 1), -0.1, 0.1) && within(K.at(2, 2), -0.1, 0.1));
    CHECK(within(b[0], 1.55, 1.75) && within(b[1], 3.25, 3.45) && within(b[2], 3.25, 3.45));
    auto u = s.solve(0.0).unwrap();
    CHECK(within(u[0], -9.1, -8.9) && within(u[1], -5.7, -5.5) && within(u[2], -2.4, -2.2));
  }
  {  // BASELINE config 1 through the window's dispatch: src/bin/static_pressure.rs:4-15, dzahui_window.rs:738-800
    auto mesh = Mesh::builder(assets + "/1dbar.obj").build_mesh_1d().unwrap();
    CHECK(mesh.node_count() == 11 && mesh.indices().size() == 60 && mesh.vertices().size() == 132);
    Solver eq = StokesParams::static_pressure().hydrostatic_pressure(100.0).density(1.0).force_function([](double) { return -10.0; }).build();
    auto solver = instantiate(eq, mesh, 350).unwrap();
    auto pr = solver->solve(1e-6).unwrap();
    CHECK(pr.size() == 11 && pr[10] == 100.0 && std::fabs(pr[0] - 109.96852572142232) < 1e-9);
    CHECK(solver->debug().rfind("StokesSolver1D", 0) == 0);
    Solver none = None{};
    CHECK(instantiate(none, mesh, 150).unwrap()->solve(0.0).unwrap().empty());
  }
  {  // the split bar through the library's communicator (one rank here): the block solve is the single-GPU solve
    auto p = DiffussionParams::time_independent().mu(1.0).b(1.0).boundary_conditions(0.0, 1.0).build();
    const size_t n = 100001;
    std::vector<double> mesh(n);
    for (size_t i = 0; i < n; ++i) mesh[i] = (double)i / (double)(n - 1);
    auto comm = Comm::init(nullptr, 0, 1).unwrap();
    CHECK(comm.world() == 1 && comm.rank() == 0);
    for (int assembly : {(int)DZB_ASSEMBLY_FAST, (int)DZB_ASSEMBLY_AUTO}) {
      dzb_options opt{assembly, DZB_SOLVE_PARALLEL, 0, 0};
      auto bar = SplitBarTimeIndependent::create(p, mesh, 150, comm, &opt).unwrap();
      CHECK(bar.rows().first == 0 && bar.rows().second == n);
      CHECK(bar.rebuild().is_ok());
      auto u = bar.solve(0.0).unwrap();
      auto whole = DiffussionSolverTimeIndependent::create(p, mesh, 150, &opt).unwrap().solve(0.0).unwrap();
      double worst = 0.0;
      for (size_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs(u[i] - whole[i]));
      CHECK(u.size() == n && u[0] == 0.0 && u[n - 1] == 1.0 && worst <= 1e-10);
    }
    CHECK(Comm::init(nullptr, 0, 2).is_err());   // more than one rank needs rank 0's id
    CHECK(Comm::init(nullptr, 1, 1).is_err());
  }
  {  // EulerSolver: the reference's integration tests (tests/euler.rs), closures given by their affine coefficients
    auto first = EulerSolver<2>::constant(1.0);  // EulerSolver::new(|_: &[f64; 2]| 1.0)
    std::array<double, 2> s2{0.0, 0.0};          // [pos, time]
    while (s2[1] <= 10.0) s2 = first.do_step(s2, 0.01);
    CHECK(s2[0] <= 10.1 && s2[0] >= 9.9);
    auto gravity = EulerSolver<3>::constant(-9.81);  // EulerSolver::new(|_: &[f64; 3]| -9.81)
    std::array<double, 3> s3{0.0, 100.0, 0.0};       // [vel, pos, time]
    while (s3[2] <= 10.0) s3 = gravity.do_step(s3, 0.01);
    CHECK(s3[0] >= -100.0 && s3[0] <= -97.0 && s3[1] >= -400.0 && s3[1] <= -389.0);
    const double lambda = std::log(2.0) / 1600.0;
    auto decay = EulerSolver<2>::affine({0.0, -lambda, 0.0});  // EulerSolver::new(|val: &[f64; 2]| -lambda * val[0])
    std::array<double, 2> q{1000.0, -1600.0};                  // [quantity, time]
    while (q[1] <= 0.0) q = decay.do_step(q, 0.5);
    CHECK(q[1] <= 0.5 && q[1] >= -0.5 && q[0] >= 490.0 && q[0] <= 510.0);
    CHECK(EulerEnsemble::create(2, 6, std::vector<double>(14, 0.0), std::vector<double>(12, 0.0)).is_err());  // no [f64; 6]
  }
}

int main(int argc, char** argv) {
  if (argc >= 3 && !std::strcmp(argv[1], "--cpu")) cpu_tests(argv[2]);
  else if (argc >= 3 && !std::strcmp(argv[1], "--gpu")) gpu_tests(argv[2]);
  else {
    std::printf("usage: host_mirror --cpu <tmp_dir> | --gpu <assets_dir>\n");
    return 2;
  }
  if (failures) {
    std::printf("%d check(s) failed\n", failures);
    return 1;
  }
  std::printf("host mirror ok\n");
  return 0;
}
