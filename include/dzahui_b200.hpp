// dzahui_b200.hpp — C++17 host side above the C ABI (dzahui_b200.h), mirroring the reference's Rust interface for the 1D
// FEM hot path name for name: parameter builders, `Error`, the `DiffEquationSolver` trait, the three solvers, `Solver`,
// `NoSolver`, `Mesh::builder(path).build_mesh_1d(mult)` and `Writer`.  The reference is a compiled (Rust) crate and Rust is
// not available in this image, so this header is the compiled-language host mirror; INTEGRATION.md holds the Rust shim.
// Header only; link with -ldzahui_b200.  Paths below are relative to /root/reference/src/lib.
//
//   Rust                                                     here
//   StokesParams::static_pressure().hydrostatic_pressure(p)  dzahui::StokesParams::static_pressure().hydrostatic_pressure(p)
//       .density(rho).force_function(Box::new(f)).build()        .density(rho).force_function(f).build()
//   XSolver::new(&params, mesh, gauss_step) -> Result<..>     XSolver::create(params, mesh, gauss_step) -> Result<XSolver>
//   solver.solve(dt) -> Result<Vec<f64>, Error>               solver.solve(dt) -> Result<std::vector<double>>
//   panic!("Params lack 'mu' term!")                          throw dzahui::Panic("Params lack 'mu' term!")
#ifndef DZAHUI_B200_HPP
#define DZAHUI_B200_HPP

#include <algorithm>
#include <array>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "dzahui_b200.h"

namespace dzahui {

// ---- error.rs:32-91: the variants the hot path can return, with the reference's Display strings -------------------
struct Error {
  enum Kind { WrongDims, Integration, PieceWiseDims, MeshParse, Io, Custom } kind;
  std::string detail;
  std::string to_string() const {
    switch (kind) {
      case WrongDims: return "One or more of the provided elements do not have the correct dimensions";
      case Integration: return "Error on integration method occurred: " + detail;
      case PieceWiseDims:
        return "Number of arguments must be one more than number of breakpoints for a piecewise function definition to make sense";
      case MeshParse: return "Unable to parse mesh file: " + detail;
      case Io: return "IO error: " + detail;
      default: return detail;
    }
  }
  static Error from_status(int rc) {
    const char* m = dzb_last_error();
    const std::string msg = m ? m : "";
    switch (rc) {
      case DZB_WRONG_DIMS: return {WrongDims, msg};
      case DZB_INTEGRATION: return {Integration, msg};
      case DZB_PIECEWISE_DIMS: return {PieceWiseDims, msg};
      case DZB_MESH_PARSE: return {MeshParse, msg};
      case DZB_IO: return {Io, msg};
      default: return {Custom, msg};  // DZB_CUDA / DZB_INVALID have no reference equivalent
    }
  }
};

// Result<T, Error>
template <class T>
class Result {
  std::variant<T, Error> v_;

 public:
  Result(T value) : v_(std::move(value)) {}
  Result(Error e) : v_(std::move(e)) {}
  bool is_ok() const { return v_.index() == 0; }
  bool is_err() const { return !is_ok(); }
  const Error& error() const { return std::get<1>(v_); }
  T& value() { return std::get<0>(v_); }
  T unwrap() {  // panics (throws) on Err like Rust's unwrap
    if (is_err()) throw std::runtime_error(error().to_string());
    return std::move(std::get<0>(v_));
  }
};

// the `panic!` of the builders' build() (diffusion_solver/mod.rs:97-181, stokes_solver/mod.rs:93-118)
struct Panic : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- stokes_solver/mod.rs:14-119, dim1.rs:26-52 ---------------------------------------------------------------------
struct StokesParams1D {
  double rho;
  double hydrostatic_pressure;
  std::function<double(double)> force_function;
};
using StaticPressureParams = StokesParams1D;

class StokesParams1DBuilder {
  std::optional<double> hydrostatic_pressure_, rho_;
  std::function<double(double)> force_function_;

 public:
  StokesParams1DBuilder hydrostatic_pressure(double v) && { hydrostatic_pressure_ = v; return std::move(*this); }
  StokesParams1DBuilder density(double v) && { rho_ = v; return std::move(*this); }
  StokesParams1DBuilder force_function(std::function<double(double)> f) && { force_function_ = std::move(f); return std::move(*this); }
  StokesParams1D build() && {
    if (!hydrostatic_pressure_) throw Panic("Params lack 'pressure' term!");
    if (!rho_) throw Panic("Params lack 'density' term!");
    if (!force_function_) throw Panic("Params lack force_function!");
    return {*rho_, *hydrostatic_pressure_, std::move(force_function_)};
  }
};
using StaticPressureParamsBuilder = StokesParams1DBuilder;

struct StokesParams {
  static StokesParams1DBuilder normal_1d() { return {}; }
  static StokesParams1DBuilder static_pressure() { return {}; }
};

// ---- diffusion_solver/mod.rs:11-181 -----------------------------------------------------------------------------------
struct DiffussionParamsTimeIndependent {
  double mu, b;
  std::array<double, 2> boundary_conditions;
};
struct DiffussionParamsTimeDependent {
  double mu, b;
  std::array<double, 2> boundary_conditions;
  std::vector<double> initial_conditions;
};

class DiffussionParamsTimeIndependentBuilder {
  std::optional<double> mu_, b_;
  std::optional<std::array<double, 2>> bc_;

 public:
  DiffussionParamsTimeIndependentBuilder mu(double v) && { mu_ = v; return std::move(*this); }
  DiffussionParamsTimeIndependentBuilder b(double v) && { b_ = v; return std::move(*this); }
  DiffussionParamsTimeIndependentBuilder boundary_conditions(double l, double r) && { bc_ = {l, r}; return std::move(*this); }
  DiffussionParamsTimeIndependent build() && {
    if (!mu_) throw Panic("Params lack 'mu' term!");
    if (!b_) throw Panic("Params lack 'b' term!");
    if (!bc_) throw Panic("Params lack boundary conditions!");
    return {*mu_, *b_, *bc_};
  }
};
class DiffussionParamsTimeDependentBuilder {
  std::optional<double> mu_, b_;
  std::optional<std::array<double, 2>> bc_;
  std::optional<std::vector<double>> ic_;

 public:
  DiffussionParamsTimeDependentBuilder mu(double v) && { mu_ = v; return std::move(*this); }
  DiffussionParamsTimeDependentBuilder b(double v) && { b_ = v; return std::move(*this); }
  DiffussionParamsTimeDependentBuilder boundary_conditions(double l, double r) && { bc_ = {l, r}; return std::move(*this); }
  DiffussionParamsTimeDependentBuilder initial_conditions(std::vector<double> ic) && { ic_ = std::move(ic); return std::move(*this); }
  DiffussionParamsTimeDependent build() && {
    if (!mu_) throw Panic("Params lack 'mu' term!");
    if (!b_) throw Panic("Params lack 'b' term!");
    if (!bc_) throw Panic("Params lack boundary conditions!");
    if (!ic_) throw Panic("Params lack initial conditions!");
    return {*mu_, *b_, *bc_, std::move(*ic_)};
  }
};
struct DiffussionParams {
  static DiffussionParamsTimeDependentBuilder time_dependent() { return {}; }
  static DiffussionParamsTimeIndependentBuilder time_independent() { return {}; }
};

// ---- mesh/mod.rs:30-68, mesh/mesh_builder.rs:204-328 ---------------------------------------------------------------
class Mesh {
  struct Deleter {
    void operator()(dzb_mesh* m) const { dzb_mesh_destroy(m); }
  };
  std::shared_ptr<dzb_mesh> h_;
  explicit Mesh(dzb_mesh* m) : h_(m, Deleter()) {}

 public:
  class Builder {
    std::string location_;

   public:
    explicit Builder(std::string location) : location_(std::move(location)) {}
    Result<Mesh> build_mesh_1d(std::optional<double> height_multiplier = std::nullopt) const {
      dzb_mesh* m = nullptr;
      const int rc = dzb_mesh_ingest_obj_1d(location_.c_str(), height_multiplier ? 1 : 0, height_multiplier.value_or(0.0), &m);
      if (rc) return Error::from_status(rc);
      return Mesh(m);
    }
  };
  static Builder builder(std::string location) { return Builder(std::move(location)); }
  static Result<Mesh> from_nodes(const std::vector<double>& nodes) {  // `mesh: Vec<f64>` handed to XSolver::new
    dzb_mesh* m = nullptr;
    const int rc = dzb_mesh_from_nodes(nodes.data(), nodes.size(), &m);
    if (rc) return Error::from_status(rc);
    return Mesh(m);
  }
  size_t node_count() const {
    size_t n = 0;
    dzb_mesh_node_count(h_.get(), &n);
    return n;
  }
  std::vector<double> filter_for_solving_1d() const {  // mesh/mod.rs:54-68
    std::vector<double> x(node_count());
    dzb_mesh_nodes(h_.get(), x.data(), x.size());
    return x;
  }
  std::vector<double> vertices() const {
    std::vector<double> v(12 * node_count());
    dzb_mesh_vertices(h_.get(), v.data(), v.size());
    return v;
  }
  std::vector<uint32_t> indices() const {
    std::vector<uint32_t> v(6 * node_count() - 6);
    dzb_mesh_indices(h_.get(), v.data(), v.size());
    return v;
  }
  // Mesh::update_gradient_1d + Drawable::get_vertices (mesh/mod.rs:73-90,108-113) on the device
  Result<std::vector<float>> update_gradient_1d(const double* device_solution) const {
    std::vector<float> v(12 * node_count());
    const int rc = dzb_mesh_update_gradient_1d(h_.get(), device_solution, node_count(), v.data(), v.size());
    if (rc) return Error::from_status(rc);
    return v;
  }
  const dzb_mesh* handle() const { return h_.get(); }
};

// ---- solvers/solver_trait.rs:12-23 -----------------------------------------------------------------------------------
class DiffEquationSolver {
 public:
  virtual ~DiffEquationSolver() = default;
  virtual Result<std::vector<double>> solve(double time_step) = 0;
  virtual std::string debug() const = 0;  // the trait's `Debug` bound
};

namespace detail {
class HandleSolver : public DiffEquationSolver {
 protected:
  struct Deleter {
    void operator()(dzb_solver* s) const { dzb_solver_destroy(s); }
  };
  std::unique_ptr<dzb_solver, Deleter> h_;
  Mesh mesh_;
  size_t n_ = 0;
  HandleSolver(dzb_solver* s, Mesh mesh) : h_(s), mesh_(std::move(mesh)) { dzb_solver_node_count(s, &n_); }

 public:
  Result<std::vector<double>> solve(double time_step) override {
    std::vector<double> out(n_);
    const int rc = dzb_solver_solve(h_.get(), time_step, out.data(), n_);
    if (rc) return Error::from_status(rc);
    return out;
  }
  // the same step, result left on the device (valid until the next call)
  Result<const double*> solve_device(double time_step) {
    const double* p = nullptr;
    const int rc = dzb_solver_solve_device(h_.get(), time_step, &p);
    if (rc) return Error::from_status(rc);
    return p;
  }
  struct Bands {
    std::vector<double> lo, di, up, rhs;
    double at(size_t i, size_t j) const { return j == i ? di[i] : (j + 1 == i ? lo[i] : (j == i + 1 ? up[i] : 0.0)); }  // [[i, j]]
  };
  Bands bands(int which) const {
    Bands b{std::vector<double>(n_), std::vector<double>(n_), std::vector<double>(n_), std::vector<double>(n_)};
    dzb_solver_bands(h_.get(), which, b.lo.data(), b.di.data(), b.up.data(), b.rhs.data(), n_);
    return b;
  }
  size_t node_count() const { return n_; }
};
inline Result<Mesh> as_mesh(const std::vector<double>& nodes) { return Mesh::from_nodes(nodes); }
}  // namespace detail

// ---- diffusion_solver/time_independent.rs:46-196 ------------------------------------------------------------------------
class DiffussionSolverTimeIndependent : public detail::HandleSolver {
  using HandleSolver::HandleSolver;

 public:
  std::array<double, 2> boundary_conditions{};
  size_t gauss_step = 0;
  double mu = 0, b = 0;
  static Result<DiffussionSolverTimeIndependent> create(const DiffussionParamsTimeIndependent& p, const Mesh& mesh, size_t gauss_step,
                                                        const dzb_options* opt = nullptr) {
    dzb_solver* s = nullptr;
    const int rc = dzb_diffusion_ti_create(mesh.handle(), p.mu, p.b, p.boundary_conditions[0], p.boundary_conditions[1], gauss_step,
                                           opt, &s);
    if (rc) return Error::from_status(rc);
    DiffussionSolverTimeIndependent out(s, mesh);
    out.boundary_conditions = p.boundary_conditions, out.gauss_step = gauss_step, out.mu = p.mu, out.b = p.b;
    return out;
  }
  static Result<DiffussionSolverTimeIndependent> create(const DiffussionParamsTimeIndependent& p, const std::vector<double>& mesh,
                                                        size_t gauss_step, const dzb_options* opt = nullptr) {
    auto m = detail::as_mesh(mesh);
    if (m.is_err()) return m.error();
    return create(p, m.value(), gauss_step, opt);
  }
  Bands stiffness_matrix() const { return bands(0); }
  std::vector<double> b_vector() const { return bands(0).rhs; }
  std::string debug() const override {
    return "DiffussionSolverTimeIndependent { boundary_conditions: [" + std::to_string(boundary_conditions[0]) + ", " +
           std::to_string(boundary_conditions[1]) + "], gauss_step: " + std::to_string(gauss_step) + ", mu: " + std::to_string(mu) +
           ", b: " + std::to_string(b) + " }";
  }
};

// ---- one bar split across the GPUs of a box: the library's communicator + the collective block solve ---------------------
// (reference: ONE solve_by_thomas over the whole system, matrix_solver/mod.rs:17-48)
class Comm {
  struct Deleter {
    void operator()(dzb_comm* c) const { dzb_comm_destroy(c); }
  };
  std::shared_ptr<dzb_comm> h_;
  int rank_ = 0, world_ = 1;
  explicit Comm(dzb_comm* c, int rank, int world) : h_(c, Deleter{}), rank_(rank), world_(world) {}

 public:
  // rank 0 makes the id and ships its 128 bytes to the other ranks (MPI, a file, a socket)
  static Result<std::array<unsigned char, DZB_COMM_ID_BYTES>> unique_id() {
    std::array<unsigned char, DZB_COMM_ID_BYTES> id{};
    const int rc = dzb_comm_unique_id(id.data());
    if (rc) return Error::from_status(rc);
    return id;
  }
  // collective; `id` may be null for a one-rank communicator
  static Result<Comm> init(const unsigned char* id, int rank, int world) {
    dzb_comm* c = nullptr;
    const int rc = dzb_comm_init(id, rank, world, &c);
    if (rc) return Error::from_status(rc);
    return Comm(c, rank, world);
  }
  dzb_comm* handle() const { return h_.get(); }
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool peer_mailboxes() const {
    int pm = 0;
    dzb_comm_info(h_.get(), nullptr, nullptr, &pm);
    return pm != 0;
  }
};

// DiffussionSolverTimeIndependent for the rows [first, last) of a bar this rank owns; `whole_mesh` is the whole bar
class SplitBarTimeIndependent {
  struct Deleter {
    void operator()(dzb_solver* s) const { dzb_solver_destroy(s); }
  };
  std::shared_ptr<dzb_solver> h_;
  Mesh mesh_;
  Comm comm_;
  size_t first_ = 0, last_ = 0;
  SplitBarTimeIndependent(dzb_solver* s, Mesh m, Comm c, size_t first, size_t last)
      : h_(s, Deleter{}), mesh_(std::move(m)), comm_(std::move(c)), first_(first), last_(last) {}

 public:
  static Result<SplitBarTimeIndependent> create(const DiffussionParamsTimeIndependent& p, const std::vector<double>& whole_mesh,
                                                size_t gauss_step, const Comm& comm, const dzb_options* opt = nullptr) {
    const size_t n = whole_mesh.size(), w = (size_t)comm.world(), r = (size_t)comm.rank();
    const size_t base = n / w, rem = n % w, first = r * base + std::min(r, rem), last = first + base + (r < rem ? 1 : 0);
    const bool hl = r > 0, hr = r + 1 < w;
    std::vector<double> nodes(whole_mesh.begin() + (first - (hl ? 1 : 0)), whole_mesh.begin() + (last + (hr ? 1 : 0)));
    auto m = Mesh::from_nodes(nodes);
    if (m.is_err()) return m.error();
    dzb_solver* s = nullptr;
    const int rc = dzb_diffusion_ti_create_block(m.value().handle(), hl, hr, p.mu, p.b, p.boundary_conditions[0], p.boundary_conditions[1],
                                                 gauss_step, opt, &s);
    if (rc) return Error::from_status(rc);
    return SplitBarTimeIndependent(s, m.value(), comm, first, last);
  }
  std::pair<size_t, size_t> rows() const { return {first_, last_}; }
  // XSolver::new again into the block's buffers (current coordinates of the block's mesh)
  Result<bool> rebuild() {
    const int rc = dzb_solver_block_rebuild(h_.get(), mesh_.handle());
    if (rc) return Error::from_status(rc);
    return true;
  }
  // collective: every rank calls it; this rank's nodal values
  Result<std::vector<double>> solve(double /*time_step*/) {
    int rc = dzb_solver_block_solve_device(h_.get(), comm_.handle(), nullptr);
    if (!rc) rc = dzb_solver_block_status(h_.get(), comm_.handle());
    if (rc) return Error::from_status(rc);
    std::vector<double> out(last_ - first_);
    rc = dzb_solver_block_read(h_.get(), out.data(), out.size());
    if (rc) return Error::from_status(rc);
    return out;
  }
};

// ---- diffusion_solver/time_dependent.rs:49-280 --------------------------------------------------------------------------
class DiffussionSolverTimeDependent : public detail::HandleSolver {
  using HandleSolver::HandleSolver;

 public:
  std::array<double, 2> boundary_conditions{};
  std::vector<double> initial_conditions;
  size_t integration_step = 0;
  double mu = 0, b = 0;
  static Result<DiffussionSolverTimeDependent> create(const DiffussionParamsTimeDependent& p, const Mesh& mesh, size_t integration_step,
                                                      const dzb_options* opt = nullptr) {
    dzb_solver* s = nullptr;
    const int rc = dzb_diffusion_td_create(mesh.handle(), p.mu, p.b, p.boundary_conditions[0], p.boundary_conditions[1],
                                           p.initial_conditions.data(), p.initial_conditions.size(), integration_step, opt, &s);
    if (rc) return Error::from_status(rc);  // WrongDims unless initial_conditions.len() == mesh.len() - 2 (:66-68)
    DiffussionSolverTimeDependent out(s, mesh);
    out.boundary_conditions = p.boundary_conditions, out.initial_conditions = p.initial_conditions;
    out.integration_step = integration_step, out.mu = p.mu, out.b = p.b;
    return out;
  }
  static Result<DiffussionSolverTimeDependent> create(const DiffussionParamsTimeDependent& p, const std::vector<double>& mesh,
                                                      size_t integration_step, const dzb_options* opt = nullptr) {
    auto m = detail::as_mesh(mesh);
    if (m.is_err()) return m.error();
    return create(p, m.value(), integration_step, opt);
  }
  Bands stiffness_matrix() const { return bands(0); }
  Bands mass_matrix() const { return bands(1); }
  std::vector<double> state() const {
    std::vector<double> u(n_);
    dzb_solver_state(h_.get(), u.data(), n_);
    return u;
  }
  std::string debug() const override {
    return "DiffussionSolverTimeDependent { integration_step: " + std::to_string(integration_step) + ", mu: " + std::to_string(mu) +
           ", b: " + std::to_string(b) + " }";
  }
};

// ---- stokes_solver/dim1.rs:66-254 -----------------------------------------------------------------------------------------
class StokesSolver1D : public detail::HandleSolver {
  using HandleSolver::HandleSolver;
  static double trampoline(double x, void* ctx) { return (*static_cast<const std::function<double(double)>*>(ctx))(x); }

 public:
  size_t gauss_step = 0;
  double hydrostatic_pressure = 0, rho = 0;
  // the closure is only borrowed for the duration of this call, like `&params` in the reference (dim1.rs:84)
  static Result<StokesSolver1D> create(const StokesParams1D& p, const Mesh& mesh, size_t gauss_step, const dzb_options* opt = nullptr) {
    dzb_solver* s = nullptr;
    const int rc = dzb_stokes1d_create(mesh.handle(), p.rho, p.hydrostatic_pressure, &trampoline,
                                       const_cast<std::function<double(double)>*>(&p.force_function), gauss_step, opt, &s);
    if (rc) return Error::from_status(rc);
    StokesSolver1D out(s, mesh);
    out.gauss_step = gauss_step, out.hydrostatic_pressure = p.hydrostatic_pressure, out.rho = p.rho;
    return out;
  }
  static Result<StokesSolver1D> create(const StokesParams1D& p, const std::vector<double>& mesh, size_t gauss_step,
                                       const dzb_options* opt = nullptr) {
    auto m = detail::as_mesh(mesh);
    if (m.is_err()) return m.error();
    return create(p, m.value(), gauss_step, opt);
  }
  Bands stiffness_matrix() const { return bands(0); }
  std::vector<double> b_vector() const { return bands(0).rhs; }
  std::string debug() const override {
    return "StokesSolver1D { gauss_step: " + std::to_string(gauss_step) + ", hydrostatic_pressure: " +
           std::to_string(hydrostatic_pressure) + ", rho: " + std::to_string(rho) + " }";
  }
};
using StaticPressureSolver = StokesSolver1D;  // stokes_solver/mod.rs:9

// ---- solvers/fem/mod.rs:27-42 and DzahuiWindow::run's dispatch (simulation/dzahui_window.rs:738-800) ---------------
class NoSolver : public DiffEquationSolver {
 public:
  Result<std::vector<double>> solve(double) override { return std::vector<double>{}; }
  std::string debug() const override { return "NoSolver"; }
};

struct None {};
using Solver = std::variant<DiffussionParamsTimeIndependent, DiffussionParamsTimeDependent, StokesParams1D, None>;

inline Result<std::unique_ptr<DiffEquationSolver>> instantiate(const Solver& solver, const Mesh& mesh, size_t integration_iteration) {
  using Ptr = std::unique_ptr<DiffEquationSolver>;
  if (auto* p = std::get_if<DiffussionParamsTimeIndependent>(&solver)) {
    auto r = DiffussionSolverTimeIndependent::create(*p, mesh, integration_iteration);
    if (r.is_err()) return r.error();
    return Ptr(new DiffussionSolverTimeIndependent(std::move(r.value())));
  }
  if (auto* p = std::get_if<DiffussionParamsTimeDependent>(&solver)) {
    auto r = DiffussionSolverTimeDependent::create(*p, mesh, integration_iteration);
    if (r.is_err()) return r.error();
    return Ptr(new DiffussionSolverTimeDependent(std::move(r.value())));
  }
  if (auto* p = std::get_if<StokesParams1D>(&solver)) {
    auto r = StokesSolver1D::create(*p, mesh, integration_iteration);
    if (r.is_err()) return r.error();
    return Ptr(new StokesSolver1D(std::move(r.value())));
  }
  return Ptr(new NoSolver());
}

// ---- quadrature/gauss_legendre.rs:5908-5927 ------------------------------------------------------------------------------
inline Result<std::pair<double, double>> quad_pair(size_t n, size_t k) {
  double theta = 0, weight = 0;
  const int rc = dzb_quad_pair(n, k, &theta, &weight);
  if (rc) return Error::from_status(rc);
  return std::make_pair(theta, weight);
}

// ---- writer.rs:27-146 -----------------------------------------------------------------------------------------------------
class Writer {
  std::string write_path_, file_prefix_;
  std::vector<std::string> variable_names_;
  Writer() = default;

 public:
  static Result<Writer> create(std::string write_path, std::string file_prefix, std::vector<std::string> variable_names,
                               bool erase_prev_dir) {
    const int rc = dzb_csv_prepare_dir(write_path.c_str(), erase_prev_dir ? 1 : 0);
    if (rc) return Error{Error::Custom, "Could not find file: Path for creating files and writing values not found"};  // Error::NotFound
    Writer w;
    w.write_path_ = std::move(write_path), w.file_prefix_ = std::move(file_prefix), w.variable_names_ = std::move(variable_names);
    return w;
  }
  Result<bool> write(double id, const std::vector<double>& vals) const {
    std::vector<const char*> names;
    for (const std::string& s : variable_names_) names.push_back(s.c_str());
    const int rc = dzb_csv_write(write_path_.c_str(), file_prefix_.c_str(), id, names.data(), names.size(), vals.data(), vals.size());
    if (rc) return Error{Error::Io, "could not write the snapshot file"};
    return true;
  }
};

// ---- EulerSolver (solvers/euler/mod.rs:26-59) ------------------------------------------------------------------------
// The reference's EulerSolver<A, F> steps one trajectory with a closure F: Fn(&A) -> f64, A = [f64; 2..5].  A host closure
// cannot run on the device: the derivative function is given by its affine coefficients instead,
//   f(values) = c[0] + c[1] * values[0] + ... + c[W] * values[W - 1],
// and many trajectories step together.  `EulerSolver<W>::affine({0, -lambda, 0}).do_step({quantity, time}, 0.5)` is the
// reference's `EulerSolver::new(|val: &[f64; 2]| -lambda * val[0]).do_step([quantity, time], 0.5)`.
class EulerEnsemble {
  dzb_euler* h_ = nullptr;
  size_t n_traj_ = 0, width_ = 0;
  EulerEnsemble() = default;

 public:
  EulerEnsemble(EulerEnsemble&& o) noexcept : h_(o.h_), n_traj_(o.n_traj_), width_(o.width_) { o.h_ = nullptr; }
  EulerEnsemble& operator=(EulerEnsemble&& o) noexcept {
    if (this != &o) {
      dzb_euler_destroy(h_);
      h_ = o.h_, n_traj_ = o.n_traj_, width_ = o.width_, o.h_ = nullptr;
    }
    return *this;
  }
  EulerEnsemble(const EulerEnsemble&) = delete;
  EulerEnsemble& operator=(const EulerEnsemble&) = delete;
  ~EulerEnsemble() { dzb_euler_destroy(h_); }
  // coeffs: [n_traj][width + 1], values: [n_traj][width], both row-major
  static Result<EulerEnsemble> create(size_t n_traj, size_t width, const std::vector<double>& coeffs, const std::vector<double>& values) {
    if (coeffs.size() != n_traj * (width + 1) || values.size() != n_traj * width) return Error{Error::WrongDims, ""};
    EulerEnsemble e;
    const int rc = dzb_euler_create(n_traj, width, coeffs.data(), values.data(), &e.h_);
    if (rc) return Error::from_status(rc);
    e.n_traj_ = n_traj, e.width_ = width;
    return e;
  }
  // `steps` x do_step(values, step) on every trajectory; returns the new values, [n_traj][width]
  Result<std::vector<double>> do_step(double step, size_t steps = 1) {
    int rc = dzb_euler_step(h_, step, steps);
    if (rc) return Error::from_status(rc);
    std::vector<double> out(n_traj_ * width_);
    rc = dzb_euler_values(h_, out.data(), out.size());
    if (rc) return Error::from_status(rc);
    return out;
  }
  Result<bool> set_values(const std::vector<double>& values) {
    const int rc = dzb_euler_set_values(h_, values.data(), values.size());
    if (rc) return Error::from_status(rc);
    return true;
  }
};

template <size_t W>
class EulerSolver {
  static_assert(W >= 2 && W <= 5, "FunctionArguments is implemented for [f64; 2] .. [f64; 5] (solvers/euler/mod.rs:19-22)");
  std::array<double, W + 1> coeffs_;
  std::unique_ptr<EulerEnsemble> ens_;
  std::array<double, W> last_{};

 public:
  explicit EulerSolver(const std::array<double, W + 1>& coeffs) : coeffs_(coeffs) {}
  static EulerSolver affine(const std::array<double, W + 1>& coeffs) { return EulerSolver(coeffs); }
  static EulerSolver constant(double value) {
    std::array<double, W + 1> c{};
    c[0] = value;
    return EulerSolver(c);
  }
  // do_step(values, step) -> next values (mod.rs:37-59); panics like the reference's `panic!("Nooo")` on failure
  std::array<double, W> do_step(const std::array<double, W>& values, double step) {
    const std::vector<double> v(values.begin(), values.end());
    if (!ens_) {
      auto r = EulerEnsemble::create(1, W, std::vector<double>(coeffs_.begin(), coeffs_.end()), v);
      if (r.is_err()) throw Panic("Nooo");
      ens_ = std::make_unique<EulerEnsemble>(std::move(r.unwrap()));
    } else if (values != last_) {
      if (ens_->set_values(v).is_err()) throw Panic("Nooo");
    }
    auto r = ens_->do_step(step);
    if (r.is_err()) throw Panic("Nooo");
    const std::vector<double> next = r.unwrap();
    std::copy(next.begin(), next.end(), last_.begin());
    return last_;
  }
};

}  // namespace dzahui
#endif  // DZAHUI_B200_HPP
