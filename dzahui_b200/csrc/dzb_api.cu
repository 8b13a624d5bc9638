// dzb_api.cu — the C ABI of include/dzahui_b200.h: handles, device memory, kernel launches.
// Build (see dzahui_b200/build.py): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
// (explicit fma() only where the FAST kernels ask for it; FAITHFUL kernels keep separate mul/add like the reference).
// No CPU fallback: every compute entry point needs a CUDA device and reports DZB_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is bound with dlopen (dzb_comm.inl), nothing links against it

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dzahui_b200.h"
#include "assembly_kernels.cuh"
#include "bigsys_kernels.cuh"
#include "euler_kernels.cuh"
#include "gl_quadrature.h"
#include "mesh2d_kernels.cuh"
#include "mesh_ingest.h"
#include "onelaunch_kernels.cuh"
#include "tridiag_kernels.cuh"

namespace {

thread_local std::string g_err;
cudaStream_t g_stream = 0;
unsigned long long g_launches = 0;
unsigned long long g_mesh_stamp = 0;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(DZB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)
#define LAUNCHED()                                                                            \
  do {                                                                                        \
    ++g_launches;                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) return fail(DZB_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

template <class T>
int dalloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) return DZB_OK;
  CK(cudaMalloc((void**)p, count * sizeof(T)));
  return DZB_OK;
}
template <class T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}
int sync_stream() {
  CK(cudaStreamSynchronize(g_stream));
  return DZB_OK;
}
inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

// ---- per-device resources (the ABI has dzb_set_device: nothing device-side may be a process-wide static) -------------
constexpr size_t SLABS = 8;  // dzb_ensemble_solve: slabs of the pipelined read-back
struct DeviceCtx {
  unsigned* d_bad = nullptr;   // rows the closed-form assembly did not cover in the last launch
  int* d_flags = nullptr;      // k_tri_classify's verdict
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t stepped[SLABS] = {}, copied = nullptr;
  bool copy_ready = false;     // set only after every create call above has succeeded
  unsigned long long* d_trace = nullptr;
};
DeviceCtx g_ctx[dzb::BS_MAX_DEVICES];
DeviceCtx& device_ctx() { return g_ctx[dzb::bigsys_device()]; }

bool env_set(const char* name) { return std::getenv(name) != nullptr; }
// environment knobs of the per-call paths, read once (knobs looked at when a handle is created are read there)
const bool g_no_multi_step = env_set("DZB_NO_MULTI_STEP"), g_no_slab_copy = env_set("DZB_NO_SLAB_COPY"), g_ol_trace = env_set("DZB_OL_TRACE");

// ---- quadrature rule cache (device copies keyed by (device, G)) -------------------------------------------------
struct DeviceRule {
  double* buf = nullptr;
  unsigned short* guess = nullptr;
  double4* uv = nullptr;
  dzb::RuleView view{};
};
std::map<std::pair<int, size_t>, DeviceRule> g_rules;

int get_rule(size_t G, dzb::RuleView* out) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  auto it = g_rules.find({dev, G});
  if (it != g_rules.end()) {
    *out = it->second.view;
    return DZB_OK;
  }
  dzb::QuadRule r;
  if (!dzb::gl_build_rule(G, &r)) return fail(DZB_INTEGRATION, "Misuse of quad_pair function: k should be smaller than n");
  const size_t m = r.m;
  std::vector<double> h;
  h.reserve(2 * m + 6 * (m + 1) + 8);
  h.insert(h.end(), r.x.begin(), r.x.end());
  h.insert(h.end(), r.w.begin(), r.w.end());
  for (const std::vector<double>* v : {&r.u0, &r.u1, &r.u2, &r.v0, &r.v1, &r.v2}) h.insert(h.end(), v->begin(), v->end());
  DeviceRule d;
  int rc = dalloc(&d.buf, h.size() + 1);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d.buf, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  double* p = d.buf;
  d.view.x = p, p += m;
  d.view.w = p, p += m;
  d.view.u0 = p, p += m + 1;
  d.view.u1 = p, p += m + 1;
  d.view.u2 = p, p += m + 1;
  d.view.v0 = p, p += m + 1;
  d.view.v1 = p, p += m + 1;
  d.view.v2 = p, p += m + 1;
  d.view.m = (int)m;
  d.view.t0 = r.t0, d.view.t1p = r.t1p, d.view.t1m = r.t1m, d.view.q = r.q;
  {  // bracket table for the kink split: guess[k] = #{j : x_j >= -1 + 2 k / GUESS_BINS} (x_j decreasing in j)
    std::vector<unsigned short> gt(dzb::GUESS_BINS + 1);
    for (int k = 0; k <= dzb::GUESS_BINS; ++k) {
      const double edge = -1.0 + 2.0 * (double)k / (double)dzb::GUESS_BINS;
      size_t c = 0;
      while (c < m && r.x[c] >= edge) ++c;
      gt[(size_t)k] = (unsigned short)std::min<size_t>(c, 65535);
    }
    CK(cudaMalloc((void**)&d.guess, gt.size() * sizeof(unsigned short)));
    CK(cudaMemcpyAsync(d.guess, gt.data(), gt.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    d.view.guess = d.guess;
    // static diffusion, lean closed form (assembly_kernels.cuh, ti_row): the same table with bit 15 set where a node lies
    // within GUESS_TOL of the bin, the four moments the row needs packed per split position, the element validity factor
    std::vector<unsigned short> gf(dzb::GUESS_BINS + 2, 0);
    for (int k = 0; k < dzb::GUESS_BINS; ++k) {
      const double lo_edge = -1.0 + 2.0 * (double)k / (double)dzb::GUESS_BINS - dzb::GUESS_TOL;
      const double hi_edge = -1.0 + 2.0 * (double)(k + 1) / (double)dzb::GUESS_BINS + dzb::GUESS_TOL;
      bool near = m > 32767;  // the count must leave bit 15 free: larger rules always probe
      for (size_t j = 0; j < m && !near; ++j) near = r.x[j] >= lo_edge && r.x[j] <= hi_edge;
      gf[(size_t)k + 1] = (unsigned short)((gt[(size_t)k + 1] & 0x7fff) | (near ? 0x8000 : 0));
    }
    unsigned short* d_gf = nullptr;
    CK(cudaMalloc((void**)&d_gf, gf.size() * sizeof(unsigned short)));
    CK(cudaMemcpyAsync(d_gf, gf.data(), gf.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, g_stream));
    std::vector<double4> uv(m + 1);
    for (size_t j = 0; j <= m; ++j) uv[j] = make_double4(r.u0[j], r.u1[j], r.v0[j], r.v1[j]);
    CK(cudaMalloc((void**)&d.uv, uv.size() * sizeof(double4)));
    CK(cudaMemcpyAsync(d.uv, uv.data(), uv.size() * sizeof(double4), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    d.view.guessf = d_gf, d.view.uv = d.uv;
    const double delta = m ? std::min(1.0 - r.x[0], 1.0 + r.x[m - 1]) : 0.0;
    d.view.kv = delta > 0.0 ? std::min(delta * 0x1p49, 0x1p30) : 0.0;
  }
  g_rules[{dev, G}] = d;
  *out = d.view;
  return DZB_OK;
}

dzb_options resolve_options(const dzb_options* opt) {
  dzb_options o{DZB_ASSEMBLY_AUTO, DZB_SOLVE_AUTO, -1, 0};
  if (opt) o = *opt;
  return o;
}
bool options_valid(const dzb_options& o) {
  return o.assembly_mode >= 0 && o.assembly_mode <= 2 && o.solve_mode >= 0 && o.solve_mode <= 2 && o.refine_steps >= -1 &&
         o.refine_steps <= 8 && (o.stored_operators == 0 || o.stored_operators == 1);
}

// ---- assembly launcher -----------------------------------------------------------------------------------------
template <int KIND>
int launch_assembly(const double* d_mesh, size_t n, size_t n_bars, size_t G, int assembly_mode, const dzb::AsmParams& P,
                    const dzb::AsmOut& o, const double* d_force) {
  dzb::RuleView rule;
  int rc = get_rule(G, &rule);
  if (rc) return rc;
  const long long total = (long long)n * (long long)n_bars;
  int mode = assembly_mode;
  if (KIND == dzb::KIND_STOKES) mode = DZB_ASSEMBLY_FAITHFUL;  // the opaque force samples make every row O(G) anyway
  // AUTO.  Static diffusion: always the reference's arithmetic.  K 1 = 0 analytically, so the stored row sums of K are
  // pure rounding noise of whoever assembled it, and on a long bar that noise IS the leading error of the solution
  // (FAST vs FAITHFUL bands agree to 1e-15 per entry, yet their exact solutions differ by 7e-5 at n = 10^6): nodal parity
  // with the reference beyond n ~ 10^3 needs its bands bit for bit.  Time dependent: the mass matrix is well conditioned
  // and S only enters scaled by dt, so the closed form holds 1e-10 (tests) and is used once the quadrature loop gets costly.
  if (mode == DZB_ASSEMBLY_AUTO) {
    if (KIND == dzb::KIND_TI) mode = DZB_ASSEMBLY_FAITHFUL;
    else mode = ((double)total * (double)(G ? G : 1) <= 16777216.0) ? DZB_ASSEMBLY_FAITHFUL : DZB_ASSEMBLY_FAST;
  }
  if (mode == DZB_ASSEMBLY_FAITHFUL) {
    const int smem_rule = (size_t)rule.m * 16 <= 40960 ? 1 : 0;
    const size_t smem = smem_rule ? (size_t)rule.m * 16 : 0;
    dzb::k_assemble_faithful<KIND><<<blocks_for(total, 128), 128, smem, g_stream>>>(d_mesh, (long long)n, (long long)n_bars,
                                                                                  rule, P, o, d_force, smem_rule);
    LAUNCHED();
  } else {
    const int smem_rule = (size_t)rule.m * 8 <= 40960 ? 1 : 0;
    const size_t smem = smem_rule ? (size_t)rule.m * 8 : 0;
    long long blocks = (total + 255) / 256;
    const long long cap = 148LL * 8 * 4;  // grid-stride: a few waves of 8 resident CTAs on each of the 148 SMs
    if (blocks > cap) blocks = cap;
    unsigned*& d_bad = device_ctx().d_bad;  // rows the closed form did not cover in the last launch (device side only)
    if (!d_bad) CK(cudaMalloc((void**)&d_bad, sizeof(unsigned)));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), g_stream));
    dzb::k_assemble_fast<KIND><<<(unsigned)blocks, 256, smem, g_stream>>>(d_mesh, (long long)n, (long long)n_bars, rule, P, o,
                                                                         smem_rule, d_bad);
    LAUNCHED();
    dzb::k_assemble_fixup<KIND><<<(unsigned)std::min<long long>(blocks, 148LL * 8), 128, 0, g_stream>>>(d_mesh, (long long)n,
                                                                                                       (long long)n_bars, rule, P, o, d_bad);
    LAUNCHED();
  }
  return DZB_OK;
}

// ---- small utility kernels ---------------------------------------------------------------------------------------
__global__ void k_td_init_state(const double* __restrict__ ic, double* __restrict__ u, long long n, long long n_bars,
                                double bc0, double bc1) {  // time_dependent.rs:71-76
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * n_bars) return;
  const long long bar = g / n, i = g - bar * n;
  u[g] = i == 0 ? bc0 : (i == n - 1 ? bc1 : ic[bar * (n - 2) + i - 1]);
}
// f64 -> f32 of a slab of states (Drawable::get_vertices' cast, mesh/mod.rs:108-113), two values per thread.  u and out are
// the same element offset into 256-byte aligned allocations: an odd offset leaves one scalar element in front of the pairs.
__global__ void __launch_bounds__(256) k_cast_f32(const double* __restrict__ u, float* __restrict__ out, long long n) {
  const long long head = (reinterpret_cast<size_t>(u) >> 3) & 1;
  if (head && blockIdx.x == 0 && threadIdx.x == 0 && n > 0) out[0] = (float)u[0];
  for (long long i = head + 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 2LL * gridDim.x * blockDim.x) {
    if (i + 1 < n) {
      const double2 v = *reinterpret_cast<const double2*>(u + i);
      *reinterpret_cast<float2*>(out + i) = make_float2((float)v.x, (float)v.y);
    } else {
      out[i] = (float)u[i];
    }
  }
}
// GL vertex / index / element tables from the sorted nodes (mesh_builder.rs:249-253, :278-307)
__global__ void k_mesh_tables(const double* __restrict__ nodes, long long n, double prom_width, double* __restrict__ vertices,
                              uint32_t* __restrict__ indices, uint32_t* __restrict__ elements) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = nodes[i];
  double* v0 = vertices + 6 * i;
  double* v1 = vertices + 6 * (n + i);
  v0[0] = x, v0[1] = 0.0, v0[2] = 0.0, v0[3] = 0.0, v0[4] = 0.0, v0[5] = 1.0;
  v1[0] = x, v1[1] = prom_width, v1[2] = 0.0, v1[3] = 0.0, v1[4] = 0.0, v1[5] = 1.0;
  const uint32_t N = (uint32_t)n, u = (uint32_t)i;
  if (n >= 2) {
    if (i == 0) {
      indices[0] = 0, indices[1] = 1, indices[2] = N;              // first triangle  :290
      indices[3] = N - 1, indices[4] = 2 * N - 1, indices[5] = 2 * N - 2;  // last triangle   :292-296
    } else if (i < n - 1) {
      uint32_t* t = indices + 6 + 6 * (i - 1);                      // :298-307
      t[0] = u, t[1] = u + N, t[2] = u + N - 1, t[3] = u, t[4] = u + 1, t[5] = u + N;
    }
    if (i < n - 1) elements[2 * i] = u, elements[2 * i + 1] = u + 1;  // element e = (node e, node e+1)
  }
}
// Mesh::update_gradient_1d (mesh/mod.rs:73-90): min/max of |u| (grid-wide partials, then one CTA), then the colour map
// + f32 cast.  mm[0] = min, mm[1] = max; part[] holds 2 doubles per block of the first pass.
__device__ __forceinline__ void block_minmax(double lo, double hi, double* out) {
  __shared__ double smin[256], smax[256];
  smin[threadIdx.x] = lo, smax[threadIdx.x] = hi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + s]);
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = smin[0], out[1] = smax[0];
}
__global__ void __launch_bounds__(256) k_absminmax(const double* __restrict__ u, long long n, double* __restrict__ part) {
  double lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = fabs(u[i]);
    lo = fmin(lo, a), hi = fmax(hi, a);
  }
  block_minmax(lo, hi, part + 2 * blockIdx.x);
}
__global__ void __launch_bounds__(256) k_absminmax_final(const double* __restrict__ part, int n_parts, double* __restrict__ mm) {
  double lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) lo = fmin(lo, part[2 * i]), hi = fmax(hi, part[2 * i + 1]);
  block_minmax(lo, hi, mm);
}
__global__ void k_gradient(const double* __restrict__ u, long long n, const double* __restrict__ mm, double* __restrict__ vertices,
                           float* __restrict__ vf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double norm_sol = (fabs(u[i]) - mm[0]) / (mm[1] - mm[0]) * (3.14159265358979323846264338327950288 / 2.);
  const double s = sin(norm_sol), c = cos(norm_sol);
  vertices[6 * i + 3] = s, vertices[6 * i + 5] = c;
  vertices[6 * (n + i) + 3] = s, vertices[6 * (n + i) + 5] = c;
  for (int k = 0; k < 6; ++k) {
    vf[6 * i + k] = (float)vertices[6 * i + k];
    vf[6 * (n + i) + k] = (float)vertices[6 * (n + i) + k];
  }
}

}  // namespace

// ================================================================================================= handles
struct dzb_mesh {
  size_t n = 0;
  bool from_obj = false;
  std::vector<double> h_nodes;  // host copy of the coordinates (may be stale after dzb_mesh_set_nodes: see host_nodes())
  bool h_stale = false;
  unsigned long long version = 0;  // process-unique stamp of the current coordinates (new at creation and at dzb_mesh_set_nodes)
  double* d_nodes = nullptr;
  double* d_vertices = nullptr;   // 12 n
  uint32_t* d_indices = nullptr;  // 6 n - 6
  uint32_t* d_elements = nullptr; // 2 (n - 1)
  float* d_vertices_f32 = nullptr;
  double* d_minmax = nullptr;
  double max_length = 0.0, prom_width = 0.0;
  float translation[3] = {0.f, 0.f, 0.f};
  // solvers that assemble on the fly from d_nodes (one-launch path): they take a private copy of the coordinates before
  // these change or go away, so a solver keeps solving the system it was built for, like the reference's owned matrices
  std::vector<dzb_solver*> dependents;
};

// ensembles of up to this many bars run the FAITHFUL kernels with a warp per bar (DZB_TW_MAX_BARS: tuning knob)
static long long tw_max_bars() {
  static const long long v = [] {
    const char* e = std::getenv("DZB_TW_MAX_BARS");
    return e ? std::atoll(e) : dzb::TW_MAX_BARS;
  }();
  return v;
}

struct dzb_euler {
  size_t T = 0, W = 0;                // trajectories, entries per state (incl. t)
  double *v = nullptr, *c = nullptr;  // SoA [W][T] state, [W + 1][T] coefficients
  // derivative function given as a postfix program (dzb_euler_create_program) instead of affine coefficients
  dzb::EulerOp* d_ops = nullptr;
  double* d_params = nullptr;         // SoA [n_params][T]
  int n_ops = 0, n_params = 0;
};

struct dzb_ensemble {
  size_t B = 0, n = 0, G = 0;
  double mu = 0, b = 0, bc0 = 0, bc1 = 0;
  bool fast = false;
  int L = 0;
  int assembly_mode = 0;
  double* d_mesh = nullptr;
  double *mlo = nullptr, *mdi = nullptr, *mup = nullptr, *slo = nullptr, *sdi = nullptr, *sup = nullptr;  // natural [B][n]
  double *c = nullptr, *den = nullptr;                                                                     // natural [B][n]
  double* rden = nullptr;  // RN(1 / den), only for the warp-per-bar FAITHFUL kernels (B <= TW_MAX_BARS)
  double *a_lo = nullptr, *a_di = nullptr, *a_up = nullptr, *alpha = nullptr, *cc = nullptr;               // interleaved [B][32 L]
  double *bdi = nullptr, *idn = nullptr, *hh = nullptr;  // lean step: unscaled diagonal, 1/den, element lengths (interleaved)
  int* d_bad = nullptr;
  int lean_L = 0, lean_wpb = 1;  // rows per lane / warps per bar of the lean layout
  bool lean_ok = false;   // the lean step applies (no row needed the quadrature-loop fallback); decided at prepare time
  dzb::LeanConsts lean_k{};
  double* u[2] = {nullptr, nullptr};
  float* u32 = nullptr;  // dzb_ensemble_solve_f32: the state as f32, what a renderer uploads (mesh/mod.rs:108-113)
  int cur = 0;
  bool prepared = false;
  double prepared_dt = 0.0;
  dzb::BigSys big;  // single long bar (n > 1024) in PARALLEL mode
  bool use_big = false, want_big = false;
};

struct dzb_solver {
  int kind = 0;  // dzb::KIND_*
  size_t n = 0;
  // static solvers
  double *lo = nullptr, *di = nullptr, *up = nullptr, *rhs = nullptr, *c = nullptr, *den = nullptr, *rden = nullptr, *out = nullptr;
  bool parallel = false;       // the partitioned solver is in use (requested and the matrix allows it)
  bool want_parallel = false;  // DZB_SOLVE_PARALLEL requested / chosen by AUTO
  bool classified_valid = false;              // the matrix class (M-matrix / dominant) was decided ...
  unsigned long long classified_version = 0;  // ... for the coordinates with this process-unique stamp (same coordinates => same bands => same class)
  bool classified_ok = false;
  int refine_steps = 0;
  dzb::BigSys big;
  // block of a longer bar (SPIKE split across GPUs): the arrays above are views shifted by `off` into allocations that
  // also hold the halo rows; block_flags = BS_OPEN_LEFT | BS_OPEN_RIGHT
  size_t off = 0, n_alloc = 0;
  int block_flags = 0;
  bool is_block = false;
  double* d_top8 = nullptr;
  // what dzb_solver_rebuild needs to assemble again
  double mu = 0, b = 0, bc0 = 0, bc1 = 0;
  size_t G = 0;
  int assembly_mode = 0;
  // time dependent: an ensemble of one bar
  dzb_ensemble* td = nullptr;
  // one-launch path (closed-form assembly fused into the partitioned solve, onelaunch_kernels.cuh): the bands are only
  // materialised when somebody asks for them (dzb_solver_bands) or the matrix turns out not to allow the partitioned solver
  bool fused = false;            // this handle may take the one-launch path
  bool bands_stale = false;      // lo/di/up/rhs do not describe the coordinates of the last rebuild
  dzb_mesh* src_mesh = nullptr;  // mesh whose d_nodes x_view points at (nullptr: x_view is the private copy x_own)
  const double* x_view = nullptr;
  double* x_own = nullptr;
  unsigned long long x_version = 0;  // coordinate stamp of x_view, for the matrix-class cache
  double* ol_red = nullptr;      // 4 x [8 OL_P1M]: the two interface rows of every CTA
  double* ol_s = nullptr;        // [n]: what pass 1 leaves per row for pass 2 (the row-sum excess)
  double* ol_pool = nullptr;               // levels B and C of the one-launch kernel: 4 + 1 arrays each (rows and their solutions)
  bool fused_bad_valid = false;            // the coordinates stamped fused_bad_version hold a row outside the closed form:
  unsigned long long fused_bad_version = 0;  // materialised bands (with the quadrature loop on such rows) for them
  int* d_cls = nullptr;          // [0] matrix-class bits of the last classifying launch, [1] != 0: a row outside the closed form
  int ol_caps[2] = {-1, -1};
  int last_path = 0;             // 0 materialised + partitioned / serial, 1 one launch from the node coordinates
  // block of a split bar on the one-launch path: x_own holds halo + own nodes behind x_pad spare entries, so that the
  // block's first OWN row is 16-byte aligned (TMA); the matrix class / closed-form coverage of new coordinates is
  // accumulated by the kernel and looked at when somebody reads the result (no host round trip inside a collective solve)
  size_t x_pad = 0;
  bool cls_pending = false;
  bool ol_checked_valid = false;             // the one-launch kernel has checked the coordinates stamped ol_checked_version itself:
  unsigned long long ol_checked_version = 0;  // later solves of the same coordinates skip the per-row checks
};

struct dzb_comm;

namespace {

int host_nodes(const dzb_mesh* cm, const std::vector<double>** out) {
  dzb_mesh* m = const_cast<dzb_mesh*>(cm);
  if (m->h_stale) {
    m->h_nodes.resize(m->n);
    CK(cudaMemcpyAsync(m->h_nodes.data(), m->d_nodes, m->n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    m->h_stale = false;
  }
  *out = &m->h_nodes;
  return DZB_OK;
}

void ensemble_free(dzb_ensemble* e) {
  if (!e) return;
  for (double** p : {&e->d_mesh, &e->mlo, &e->mdi, &e->mup, &e->slo, &e->sdi, &e->sup, &e->c, &e->den, &e->rden, &e->a_lo, &e->a_di,
                     &e->a_up, &e->alpha, &e->cc, &e->u[0], &e->u[1], &e->bdi, &e->idn, &e->hh})
    dfree(*p);
  if (e->d_bad) cudaFree(e->d_bad), e->d_bad = nullptr;
  if (e->u32) cudaFree(e->u32), e->u32 = nullptr;
  dzb::bigsys_free(&e->big);
  delete e;
}

// ---- one-launch path: a solver that assembles on the fly reads the mesh's coordinates in place -----------------------
void solver_detach(dzb_solver* s) {
  if (!s->src_mesh) return;
  std::vector<dzb_solver*>& d = s->src_mesh->dependents;
  d.erase(std::remove(d.begin(), d.end(), s), d.end());
  s->src_mesh = nullptr;
}
void solver_attach(dzb_solver* s, const dzb_mesh* cm) {
  dzb_mesh* m = const_cast<dzb_mesh*>(cm);
  if (s->src_mesh != m) {
    solver_detach(s);
    m->dependents.push_back(s);
    s->src_mesh = m;
  }
  s->x_view = m->d_nodes, s->x_version = m->version;
}
// m's coordinates are about to change or go away: every dependent keeps a private copy of the ones it was built for
int mesh_release_dependents(dzb_mesh* m) {
  int rc = DZB_OK;
  for (dzb_solver* s : m->dependents) {
    s->src_mesh = nullptr;
    if (!s->x_own && (rc = dalloc(&s->x_own, s->n))) {
      s->x_view = nullptr;
      continue;
    }
    if (cudaMemcpyAsync(s->x_own, m->d_nodes, s->n * sizeof(double), cudaMemcpyDeviceToDevice, g_stream) != cudaSuccess)
      rc = fail(DZB_CUDA, "D2D node coordinates");
    s->x_view = s->x_own;
  }
  m->dependents.clear();
  return rc;
}

template <int L>
int launch_prepare(dzb_ensemble* e, double dt) {
  long long blocks = (long long)e->B;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  dzb::k_td_prepare<L><<<(unsigned)blocks, 256, 0, g_stream>>>(e->mlo, e->mdi, e->mup, e->slo, e->sdi, e->sup, e->c, e->den, e->a_lo,
                                                               e->a_di, e->a_up, e->alpha, e->cc, (int)e->n, (long long)e->B, dt);
  LAUNCHED();
  return DZB_OK;
}
template <int L>
int launch_step_fast(dzb_ensemble* e) {
  constexpr int WARPS = 4;
  constexpr size_t smem = WARPS * (32 * L + 34) * sizeof(double);
  if (smem > 48 * 1024) {
    static dzb::PerDevice<char> attr;
    bool fresh;
    attr.slot(&fresh);
    if (fresh && cudaFuncSetAttribute(dzb::k_td_step_fast<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      attr.forget();
      return fail(DZB_CUDA, "cudaFuncSetAttribute(k_td_step_fast)");
    }
  }
  const unsigned blocks = (unsigned)((e->B + WARPS - 1) / WARPS);
  dzb::k_td_step_fast<L, WARPS><<<blocks, 32 * WARPS, smem, g_stream>>>(e->u[e->cur], e->u[e->cur ^ 1], e->a_lo, e->a_di, e->a_up,
                                                                         e->alpha, e->cc, (int)e->n, (long long)e->B, e->bc0, e->bc1);
  LAUNCHED();
  return DZB_OK;
}
template <int L, int WPB>
int launch_prepare_lean(dzb_ensemble* e, double dt, double x_hi, double x_lo) {
  long long blocks = (long long)e->B;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  dzb::k_td_prepare_lean<L, WPB><<<(unsigned)blocks, 256, 0, g_stream>>>(e->d_mesh, e->mdi, e->sdi, e->den, e->bdi, e->idn, e->hh,
                                                                         (int)e->n, (long long)e->B, dt, x_hi, x_lo, e->d_bad);
  LAUNCHED();
  return DZB_OK;
}
template <int L, int WPB>
int launch_step_lean(dzb_ensemble* e) {
  dzb::k_td_step_lean<L, WPB><<<(unsigned)e->B, 32 * WPB, 0, g_stream>>>(e->u[e->cur], e->u[e->cur ^ 1], e->bdi, e->idn, e->hh, (int)e->n,
                                                                          e->bc0, e->bc1, e->lean_k);
  LAUNCHED();
  return DZB_OK;
}
// the same step on bars [bar0, bar0 + nb) only (every per-bar array is bar-major), for the pipelined read-back
template <int L, int WPB>
int launch_step_lean_slab(dzb_ensemble* e, size_t bar0, size_t nb) {
  const size_t uo = bar0 * e->n, po = bar0 * (size_t)(32 * WPB * L);
  dzb::k_td_step_lean<L, WPB><<<(unsigned)nb, 32 * WPB, 0, g_stream>>>(e->u[e->cur] + uo, e->u[e->cur ^ 1] + uo, e->bdi + po, e->idn + po,
                                                                      e->hh + po, (int)e->n, e->bc0, e->bc1, e->lean_k);
  LAUNCHED();
  return DZB_OK;
}
template <int L, int WPB>
int launch_step_lean_multi(dzb_ensemble* e, int steps) {
  dzb::k_td_step_lean_multi<L, WPB><<<(unsigned)e->B, 32 * WPB, 0, g_stream>>>(e->u[e->cur], e->u[e->cur ^ 1], e->bdi, e->idn, e->hh,
                                                                                (int)e->n, e->bc0, e->bc1, e->lean_k, steps);
  LAUNCHED();
  return DZB_OK;
}
// lean layout: (rows per lane, warps per bar)
#define DISPATCH_LEAN(E_, FN, ...)                                \
  switch ((E_)->lean_L * 8 + (E_)->lean_wpb) {                    \
    case 1 * 8 + 1: rc = FN<1, 1>(__VA_ARGS__); break;            \
    case 2 * 8 + 1: rc = FN<2, 1>(__VA_ARGS__); break;            \
    case 4 * 8 + 1: rc = FN<4, 1>(__VA_ARGS__); break;            \
    case 8 * 8 + 1: rc = FN<8, 1>(__VA_ARGS__); break;            \
    case 8 * 8 + 2: rc = FN<8, 2>(__VA_ARGS__); break;            \
    case 8 * 8 + 4: rc = FN<8, 4>(__VA_ARGS__); break;            \
    case 16 * 8 + 2: rc = FN<16, 2>(__VA_ARGS__); break;          \
    default: rc = fail(DZB_INVALID, "bad lean layout");           \
  }
#define DISPATCH_L(L_, FN, ...)                  \
  switch (L_) {                                  \
    case 1: rc = FN<1>(__VA_ARGS__); break;      \
    case 2: rc = FN<2>(__VA_ARGS__); break;      \
    case 4: rc = FN<4>(__VA_ARGS__); break;      \
    case 8: rc = FN<8>(__VA_ARGS__); break;      \
    case 16: rc = FN<16>(__VA_ARGS__); break;    \
    case 32: rc = FN<32>(__VA_ARGS__); break;    \
    default: rc = fail(DZB_INVALID, "bad chunk length"); \
  }

int ensemble_step_once(dzb_ensemble* e, double dt) {
  int rc = DZB_OK;
  if (e->use_big) {
    rc = dzb::bigsys_td_step(&e->big, e->u[e->cur], e->u[e->cur ^ 1], e->mlo, e->mdi, e->mup, e->slo, e->sdi, e->sup, dt, e->bc0,
                             e->bc1, g_stream, &g_launches);
    if (rc) return fail(DZB_CUDA, dzb::bigsys_error());
  } else if (e->fast) {
    if (!e->prepared || e->prepared_dt != dt) {
      if (e->lean_ok) {  // operator rebuilt on chip from (h, 1/den, unscaled diagonal): 40 B per DOF-step
        const size_t np = e->B * 32 * (size_t)e->lean_wpb * (size_t)e->lean_L;
        for (double** p : {&e->bdi, &e->idn, &e->hh})
          if (!*p && (rc = dalloc(p, np))) return rc;
        if (!e->d_bad) CK(cudaMalloc((void**)&e->d_bad, sizeof(int)));
        CK(cudaMemsetAsync(e->d_bad, 0, sizeof(int), g_stream));
        dzb::QuadRule rule;
        if (!dzb::gl_build_rule(e->G, &rule) || rule.m == 0) e->lean_ok = false;
        else {
          DISPATCH_LEAN(e, launch_prepare_lean, e, dt, rule.x[0], rule.x[rule.m - 1]);
          if (rc) return rc;
          int bad = 0;
          CK(cudaMemcpyAsync(&bad, e->d_bad, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
          CK(cudaStreamSynchronize(g_stream));
          if (bad) e->lean_ok = false;  // some row needed the quadrature-loop fallback: use the stored operators
          e->lean_k = dzb::LeanConsts{dt * e->mu * rule.t0 / 2.0, rule.q / 8.0, dt * e->b * rule.t1p / 4.0, dt * e->b * rule.t1m / 4.0};
        }
      }
      if (!e->lean_ok) {
        const size_t np = e->B * 32 * (size_t)e->L;
        for (double** p : {&e->a_lo, &e->a_di, &e->a_up, &e->alpha, &e->cc})
          if (!*p && (rc = dalloc(p, np))) return rc;
        DISPATCH_L(e->L, launch_prepare, e, dt);
        if (rc) return rc;
      }
      e->prepared = true;
      e->prepared_dt = dt;
    }
    if (e->lean_ok) {
      DISPATCH_LEAN(e, launch_step_lean, e);
    } else {
      DISPATCH_L(e->L, launch_step_fast, e);
    }
    if (rc) return rc;
  } else {
    if ((long long)e->B <= tw_max_bars())  // few bars: a warp per bar (same arithmetic, operands staged through shared memory)
      dzb::k_td_step_faithful_warp<<<(unsigned)e->B, 32, 0, g_stream>>>(e->u[e->cur], e->u[e->cur ^ 1], e->mlo, e->mdi, e->mup, e->slo,
                                                                        e->sdi, e->sup, e->c, e->den, e->rden, (long long)e->n, dt, e->bc0, e->bc1);
    else
      dzb::k_td_step_faithful<<<blocks_for((long long)e->B, 64), 64, 0, g_stream>>>(e->u[e->cur], e->u[e->cur ^ 1], e->mlo, e->mdi,
                                                                                   e->mup, e->slo, e->sdi, e->sup, e->c, e->den,
                                                                                   (long long)e->n, (long long)e->B, dt, e->bc0, e->bc1);
    LAUNCHED();
  }
  e->cur ^= 1;
  return DZB_OK;
}

// true when the partitioned solver is reliable on (lo, di, up): an M-matrix or a strictly dominant positive one
int tri_parallel_ok(const double* lo, const double* di, const double* up, size_t n, bool* ok, int open_ends = 0) {
  int*& d_flags = device_ctx().d_flags;
  if (!d_flags) CK(cudaMalloc((void**)&d_flags, sizeof(int)));
  CK(cudaMemsetAsync(d_flags, 0, sizeof(int), g_stream));
  dzb::k_tri_classify<<<dzb::bigsys_grid(n), 256, 0, g_stream>>>(lo, di, up, (long long)n, d_flags, open_ends);
  LAUNCHED();
  int f = 0;
  CK(cudaMemcpyAsync(&f, d_flags, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  *ok = f != 3;  // at least one of the two classes holds on every interior row
  return DZB_OK;
}
// assembly of M and S plus the per-matrix Thomas factors, into the buffers the handle already owns
int ensemble_assemble(dzb_ensemble* e) {
  dzb::AsmParams P{e->mu, e->b, e->bc0, e->bc1, 0.0, 0.0};
  dzb::AsmOut o{e->slo, e->sdi, e->sup, nullptr, e->mlo, e->mdi, e->mup};
  int rc = launch_assembly<dzb::KIND_TD>(e->d_mesh, e->n, e->B, e->G, e->assembly_mode, P, o, nullptr);
  if (rc) return rc;
  if (e->want_big) {  // one long bar: the partitioned solver if the mass matrix allows it (it does for linear hats)
    bool ok = false;
    if ((rc = tri_parallel_ok(e->mlo, e->mdi, e->mup, e->n, &ok))) return rc;
    e->use_big = ok;
  }
  if (!e->use_big) {
    if ((long long)e->B <= tw_max_bars() && !e->rden && (rc = dalloc(&e->rden, e->B * e->n))) return rc;
    if ((long long)e->B <= tw_max_bars())
      dzb::k_thomas_factor_warp<<<(unsigned)e->B, 32, 0, g_stream>>>(e->mlo, e->mdi, e->mup, e->c, e->den, e->rden, (long long)e->n);
    else {  // many bars: a warp per 32 bars over padded shared-memory tiles (coalesced; same arithmetic as k_thomas_factor)
      static dzb::PerDevice<char> attr;
      bool fresh;
      attr.slot(&fresh);
      if (fresh && cudaFuncSetAttribute(dzb::k_thomas_factor_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dzb::TF_SMEM) != cudaSuccess) {
        attr.forget();
        return fail(DZB_CUDA, "cudaFuncSetAttribute(k_thomas_factor_tile)");
      }
      dzb::k_thomas_factor_tile<<<blocks_for((long long)e->B, 32 * dzb::TF_WARPS), 32 * dzb::TF_WARPS, dzb::TF_SMEM, g_stream>>>(
          e->mlo, e->mdi, e->mup, e->c, e->den, (long long)e->n, (long long)e->B);
    }
    LAUNCHED();
  }
  e->prepared = false;
  return DZB_OK;
}

int ensemble_create_impl(const double* h_meshes, size_t B, size_t n, double mu, double b, double bc0, double bc1, const double* h_ic,
                         size_t n_ic, size_t G, const dzb_options* opt_in, dzb_ensemble** out) {
  if (!out) return fail(DZB_INVALID, "null output handle");
  *out = nullptr;
  if (!h_meshes || B == 0) return fail(DZB_INVALID, "null meshes / empty ensemble");
  const dzb_options opt = resolve_options(opt_in);
  if (!options_valid(opt)) return fail(DZB_INVALID, "bad dzb_options");
  if (n < 3) return fail(DZB_WRONG_DIMS, "a bar needs at least 3 nodes");
  if (n_ic != n - 2 || (!h_ic && n_ic)) return fail(DZB_WRONG_DIMS, "initial_conditions.len() != mesh.len() - 2");  // time_dependent.rs:66-68
  dzb::RuleView rv;
  int rc = get_rule(G, &rv);  // surfaces the reference's Integration error before any allocation
  if (rc) return rc;

  dzb_ensemble* e = new dzb_ensemble;
  e->B = B, e->n = n, e->G = G, e->mu = mu, e->b = b, e->bc0 = bc0, e->bc1 = bc1;
  const size_t tot = B * n;
  auto bail = [&](int code) {
    ensemble_free(e);
    return code;
  };
  for (double** p : {&e->d_mesh, &e->mlo, &e->mdi, &e->mup, &e->slo, &e->sdi, &e->sup, &e->c, &e->den, &e->u[0], &e->u[1]})
    if ((rc = dalloc(p, tot))) return bail(rc);
  if (cudaMemcpyAsync(e->d_mesh, h_meshes, tot * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
    return bail(fail(DZB_CUDA, "H2D meshes"));
  {  // initial state = [bc0, ic..., bc1]; the IC buffer is staged through u[1]
    if (n_ic) {
      if (cudaMemcpyAsync(e->u[1], h_ic, B * n_ic * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
        return bail(fail(DZB_CUDA, "H2D initial conditions"));
    }
    k_td_init_state<<<blocks_for((long long)tot, 256), 256, 0, g_stream>>>(e->u[1], e->u[0], (long long)n, (long long)B, bc0, bc1);
    ++g_launches;
  }
  int solve_mode = opt.solve_mode;
  if (solve_mode == DZB_SOLVE_AUTO)  // many long bars: the per-bar kernels only cover 1024 nodes, the serial one covers anything
    solve_mode = (tot <= 65536 || (n > 1024 && B != 1)) ? DZB_SOLVE_FAITHFUL : DZB_SOLVE_PARALLEL;
  if (solve_mode == DZB_SOLVE_PARALLEL && n > 1024 && B != 1)
    return bail(fail(DZB_INVALID, "PARALLEL ensembles support bars of up to 1024 nodes (longer bars: one bar per handle)"));
  e->assembly_mode = opt.assembly_mode;
  if (solve_mode == DZB_SOLVE_PARALLEL && n > 1024) {
    e->want_big = true;
    if (dzb::bigsys_prepare(&e->big, n)) return bail(fail(DZB_CUDA, dzb::bigsys_error()));
  } else if (solve_mode == DZB_SOLVE_PARALLEL) {
    int L = 1;
    while ((size_t)32 * L < n) L <<= 1;
    e->L = L;
    e->fast = true;
    e->lean_ok = opt.stored_operators == 0;  // dzb_options.stored_operators = 1 keeps the pre-scaled operator arrays (k_td_step_fast)
    // at most 8 rows per lane (~100 registers): 1 warp per bar up to 256 nodes, 2 up to 512, 4 up to 1024
    e->lean_wpb = n > 512 ? 4 : (n > 256 ? 2 : 1);
    if (const char* w = std::getenv("DZB_LEAN_WPB")) {  // tuning knob: 2 warps x 16 rows for 513..1024 nodes
      if (std::atoi(w) == 2 && n > 512) e->lean_wpb = 2;
    }
    e->lean_L = 1;
    while ((size_t)32 * e->lean_wpb * e->lean_L < n) e->lean_L <<= 1;
  }
  if ((rc = ensemble_assemble(e))) return bail(rc);
  if (cudaStreamSynchronize(g_stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
    return bail(fail(DZB_CUDA, std::string("ensemble setup: ") + cudaGetErrorString(cudaGetLastError())));
  *out = e;
  return DZB_OK;
}

}  // namespace

// ================================================================================================= C ABI
extern "C" {

const char* dzb_last_error(void) { return g_err.c_str(); }
int dzb_version(void) { return 100; }

int dzb_device_count(int* count) {
  if (!count) return fail(DZB_INVALID, "null pointer");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(DZB_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  return DZB_OK;
}
int dzb_set_device(int device) {
  CK(cudaSetDevice(device));
  return DZB_OK;
}
int dzb_set_stream(void* s) {
  g_stream = (cudaStream_t)s;
  return DZB_OK;
}
int dzb_synchronize(void) { return sync_stream(); }
// Page-locked host memory for result buffers: a D2H copy into it runs at PCIe speed and asynchronously, into pageable
// memory (a plain Vec<f64>) the driver stages it at a fraction of that.
int dzb_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(DZB_INVALID, "null pointer");
  *ptr = nullptr;
  CK(cudaMallocHost(ptr, bytes ? bytes : 1));
  return DZB_OK;
}
int dzb_host_free(void* ptr) {
  if (ptr) CK(cudaFreeHost(ptr));
  return DZB_OK;
}
int dzb_launch_count(unsigned long long* count, int reset) {
  if (count) *count = g_launches;
  if (reset) g_launches = 0;
  return DZB_OK;
}

// ------------------------------------------------------------------------------------------------- quadrature
int dzb_quad_pair(size_t n, size_t k, double* theta, double* weight) {
  if (!theta || !weight) return fail(DZB_INVALID, "null pointer");
  if (!dzb::gl_quad_pair(n, k, theta, weight))
    return fail(DZB_INTEGRATION, "Misuse of quad_pair function: k should be smaller than n");
  return DZB_OK;
}
int dzb_quad_table(size_t G, double* x, double* w) {
  dzb::QuadRule r;
  if (!dzb::gl_build_rule(G, &r)) return fail(DZB_INTEGRATION, "Misuse of quad_pair function: k should be smaller than n");
  if (r.m && (!x || !w)) return fail(DZB_INVALID, "null pointer");
  std::memcpy(x, r.x.data(), r.m * sizeof(double));
  std::memcpy(w, r.w.data(), r.m * sizeof(double));
  return DZB_OK;
}

// ------------------------------------------------------------------------------------------------- mesh
static int mesh_finish(dzb_mesh* m, bool tables) {
  const size_t n = m->n;
  m->version = ++g_mesh_stamp;
  int rc;
  if ((rc = dalloc(&m->d_nodes, n))) return rc;
  CK(cudaMemcpyAsync(m->d_nodes, m->h_nodes.data(), n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  if (tables) {
    if ((rc = dalloc(&m->d_vertices, 12 * n))) return rc;
    if ((rc = dalloc(&m->d_indices, n > 1 ? 6 * n - 6 : 6))) return rc;
    if ((rc = dalloc(&m->d_elements, n > 1 ? 2 * (n - 1) : 2))) return rc;
    k_mesh_tables<<<blocks_for((long long)n, 256), 256, 0, g_stream>>>(m->d_nodes, (long long)n, m->prom_width, m->d_vertices,
                                                                       m->d_indices, m->d_elements);
    LAUNCHED();
  }
  return sync_stream();
}

int dzb_mesh_ingest_obj_1d(const char* path, int has_mult, double mult, dzb_mesh** out) {
  if (!out || !path) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  dzb::Ingested1D ing;
  const dzb::IngestStatus st = dzb::ingest_obj_1d(path, has_mult != 0, mult, &ing);
  if (st == dzb::IngestStatus::Io) return fail(DZB_IO, ing.message);
  if (st == dzb::IngestStatus::Parse) return fail(DZB_MESH_PARSE, ing.message);
  dzb_mesh* m = new dzb_mesh;
  m->n = ing.nodes.size();
  m->from_obj = true;
  m->h_nodes.swap(ing.nodes);
  m->max_length = ing.max_length, m->prom_width = ing.prom_width;
  for (int k = 0; k < 3; ++k) m->translation[k] = ing.translation[k];
  int rc = mesh_finish(m, true);
  if (rc) {
    dzb_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return DZB_OK;
}
int dzb_mesh_from_nodes(const double* x, size_t n, dzb_mesh** out) {
  if (!out || !x) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  if (n < 1) return fail(DZB_WRONG_DIMS, "empty mesh");
  dzb_mesh* m = new dzb_mesh;
  m->n = n;
  m->h_nodes.assign(x, x + n);
  m->max_length = -x[0] + x[n - 1];
  int rc = mesh_finish(m, false);
  if (rc) {
    dzb_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return DZB_OK;
}
int dzb_mesh_set_nodes(dzb_mesh* m, const double* x, size_t n) {
  if (!m || !x) return fail(DZB_INVALID, "null pointer");
  if (n != m->n) return fail(DZB_WRONG_DIMS, "node count");
  if (m->from_obj) return fail(DZB_INVALID, "only meshes made by dzb_mesh_from_nodes can be re-pointed");
  if (!m->dependents.empty()) {
    int rc = mesh_release_dependents(m);
    if (rc) return rc;
  }
  m->h_stale = true;  // refreshed from the device only if a later call needs host coordinates
  m->version = ++g_mesh_stamp;
  m->max_length = -x[0] + x[n - 1];
  CK(cudaMemcpyAsync(m->d_nodes, x, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  return sync_stream();
}
int dzb_mesh_node_count(const dzb_mesh* m, size_t* n) {
  if (!m || !n) return fail(DZB_INVALID, "null pointer");
  *n = m->n;
  return DZB_OK;
}
int dzb_mesh_nodes(const dzb_mesh* m, double* out, size_t n) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (n != m->n) return fail(DZB_WRONG_DIMS, "node buffer length");
  CK(cudaMemcpyAsync(out, m->d_nodes, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh_vertices(const dzb_mesh* m, double* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (!m->d_vertices || len != 12 * m->n) return fail(DZB_WRONG_DIMS, "vertex buffer length (12 n; only .obj meshes carry vertices)");
  CK(cudaMemcpyAsync(out, m->d_vertices, len * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh_indices(const dzb_mesh* m, uint32_t* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (!m->d_indices || m->n < 2 || len != 6 * m->n - 6) return fail(DZB_WRONG_DIMS, "index buffer length (6 n - 6)");
  CK(cudaMemcpyAsync(out, m->d_indices, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh_elements(const dzb_mesh* m, uint32_t* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (!m->d_elements || m->n < 2 || len != 2 * (m->n - 1)) return fail(DZB_WRONG_DIMS, "element buffer length (2 (n - 1))");
  CK(cudaMemcpyAsync(out, m->d_elements, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh_info(const dzb_mesh* m, double* max_length, double* prom_width, float translation[3]) {
  if (!m) return fail(DZB_INVALID, "null pointer");
  if (max_length) *max_length = m->max_length;
  if (prom_width) *prom_width = m->prom_width;
  if (translation)
    for (int k = 0; k < 3; ++k) translation[k] = m->translation[k];
  return DZB_OK;
}
int dzb_mesh_device_nodes(const dzb_mesh* m, const double** p) {
  if (!m || !p) return fail(DZB_INVALID, "null pointer");
  *p = m->d_nodes;
  return DZB_OK;
}
int dzb_mesh_update_gradient_1d(dzb_mesh* m, const double* d_solution, size_t n, float* out, size_t len) {
  if (!m || !d_solution || !out) return fail(DZB_INVALID, "null pointer");
  if (!m->d_vertices || n != m->n || len != 12 * m->n) return fail(DZB_WRONG_DIMS, "gradient buffers");
  int rc;
  if (!m->d_vertices_f32 && (rc = dalloc(&m->d_vertices_f32, 12 * n))) return rc;
  constexpr int MAX_PARTS = 148 * 4;
  if (!m->d_minmax && (rc = dalloc(&m->d_minmax, 2 + 2 * MAX_PARTS))) return rc;
  const int parts = (int)std::min<long long>(MAX_PARTS, ((long long)n + 255) / 256);
  k_absminmax<<<parts, 256, 0, g_stream>>>(d_solution, (long long)n, m->d_minmax + 2);
  LAUNCHED();
  k_absminmax_final<<<1, 256, 0, g_stream>>>(m->d_minmax + 2, parts, m->d_minmax);
  LAUNCHED();
  k_gradient<<<blocks_for((long long)n, 256), 256, 0, g_stream>>>(d_solution, (long long)n, m->d_minmax, m->d_vertices, m->d_vertices_f32);
  LAUNCHED();
  CK(cudaMemcpyAsync(out, m->d_vertices_f32, len * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
void dzb_mesh_destroy(dzb_mesh* m) {
  if (!m) return;
  if (!m->dependents.empty()) {
    mesh_release_dependents(m);
    cudaStreamSynchronize(g_stream);
  }
  dfree(m->d_nodes), dfree(m->d_vertices), dfree(m->d_indices), dfree(m->d_elements), dfree(m->d_vertices_f32), dfree(m->d_minmax);
  delete m;
}

// ------------------------------------------------------------------------------------------------- 2D meshes (f3)
struct dzb_mesh2d {
  size_t n_vertices = 0, n_triangles = 0, n_boundary = 0;
  double* d_vertices = nullptr;   // 6 V
  uint32_t* d_indices = nullptr;  // 3 T
  uint32_t* d_boundary = nullptr; // n_boundary, ascending
  double max_length = 0.0;
  float translation[3] = {0.f, 0.f, 0.f};
};

// boundary vertices of a device-resident triangle list; d_out needs room for n_vertices entries
static int boundary_vertices_device(const uint32_t* d_tri, size_t n_tri, size_t n_vert, uint32_t* d_out, size_t* n_out) {
  *n_out = 0;
  if (n_tri == 0 || n_vert == 0) return DZB_OK;
  if (n_vert > 0xfffffffeull) return fail(DZB_WRONG_DIMS, "more vertices than a u32 index can address");  // (0xffffffff, 0xffffffff) is the empty-slot key
  unsigned long long slots = 1024;
  while (slots < 8ull * n_tri) slots <<= 1;  // 3 T edges at most -> load factor <= 0.375
  unsigned long long* keys = nullptr;
  unsigned *cnt = nullptr, *blk = nullptr;
  unsigned char* flag = nullptr;
  unsigned long long* d_total = nullptr;
  const size_t nblk = (n_vert + 1023) / 1024;
  int rc = DZB_OK;
  auto cleanup = [&]() {
    for (void* p : {(void*)keys, (void*)cnt, (void*)blk, (void*)flag, (void*)d_total})
      if (p) cudaFree(p);
  };
#define CK2(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      cleanup();                                                                                    \
      return fail(DZB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    }                                                                                               \
  } while (0)
  CK2(cudaMalloc((void**)&keys, slots * sizeof(unsigned long long)));
  CK2(cudaMalloc((void**)&cnt, slots * sizeof(unsigned)));
  CK2(cudaMalloc((void**)&flag, n_vert));
  CK2(cudaMalloc((void**)&blk, nblk * sizeof(unsigned)));
  CK2(cudaMalloc((void**)&d_total, 2 * sizeof(unsigned long long)));  // [0] boundary vertices, [1] edges with an index out of range
  CK2(cudaMemsetAsync(d_total, 0, 2 * sizeof(unsigned long long), g_stream));
  CK2(cudaMemsetAsync(keys, 0xff, slots * sizeof(unsigned long long), g_stream));
  CK2(cudaMemsetAsync(cnt, 0, slots * sizeof(unsigned), g_stream));
  CK2(cudaMemsetAsync(flag, 0, n_vert, g_stream));
  const unsigned grid_e = (unsigned)std::min<unsigned long long>((3ull * n_tri + 255) / 256, 148ull * 32);
  dzb::k_edge_count<<<grid_e, 256, 0, g_stream>>>(d_tri, (long long)n_tri, (uint32_t)n_vert, keys, cnt, slots - 1, d_total + 1);
  ++g_launches;
  const unsigned grid_s = (unsigned)std::min<unsigned long long>((slots + 255) / 256, 148ull * 32);
  dzb::k_edge_flag<<<grid_s, 256, 0, g_stream>>>(keys, cnt, slots, flag);
  ++g_launches;
  dzb::k_flag_count<<<(unsigned)nblk, 256, 0, g_stream>>>(flag, (long long)n_vert, blk);
  ++g_launches;
  dzb::k_block_scan<<<1, 1024, 0, g_stream>>>(blk, (long long)nblk, d_total);
  ++g_launches;
  dzb::k_flag_scatter<<<(unsigned)nblk, 256, 0, g_stream>>>(flag, (long long)n_vert, blk, d_out);
  ++g_launches;
  unsigned long long total[2] = {0, 0};
  CK2(cudaMemcpyAsync(total, d_total, sizeof total, cudaMemcpyDeviceToHost, g_stream));
  CK2(cudaStreamSynchronize(g_stream));
  CK2(cudaGetLastError());
#undef CK2
  cleanup();
  if (total[1]) return fail(DZB_MESH_PARSE, "a triangle refers to a vertex that does not exist");
  *n_out = (size_t)total[0];
  return rc;
}

int dzb_mesh_ingest_obj_2d(const char* path, dzb_mesh2d** out) {
  if (!out || !path) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  dzb::Ingested2D ing;
  const dzb::IngestStatus st = dzb::ingest_obj_2d(path, &ing);
  if (st == dzb::IngestStatus::Io) return fail(DZB_IO, ing.message);
  if (st == dzb::IngestStatus::Parse) return fail(DZB_MESH_PARSE, ing.message);
  const size_t V = ing.xy.size() / 2, T = ing.triangles.size() / 3;
  for (unsigned v : ing.triangles)
    if (v >= V) return fail(DZB_MESH_PARSE, "a face refers to a vertex that does not exist");
  dzb_mesh2d* m = new dzb_mesh2d;
  m->n_vertices = V, m->n_triangles = T, m->max_length = ing.max_length;
  for (int k = 0; k < 3; ++k) m->translation[k] = ing.translation[k];
  double* d_xy = nullptr;
  int rc = dalloc(&m->d_vertices, 6 * V);
  if (!rc) rc = dalloc(&m->d_indices, 3 * T);
  if (!rc) rc = dalloc(&m->d_boundary, V);
  if (!rc) rc = dalloc(&d_xy, 2 * V);
  if (!rc && V) {
    if (cudaMemcpyAsync(d_xy, ing.xy.data(), 2 * V * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
        (T && cudaMemcpyAsync(m->d_indices, ing.triangles.data(), 3 * T * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream) !=
                  cudaSuccess))
      rc = fail(DZB_CUDA, "H2D mesh");
  }
  if (!rc && V) {
    dzb::k_mesh2d_vertices<<<blocks_for((long long)V, 256), 256, 0, g_stream>>>(d_xy, (long long)V, m->d_vertices);
    ++g_launches;
  }
  if (!rc) rc = boundary_vertices_device(m->d_indices, T, V, m->d_boundary, &m->n_boundary);
  if (!rc && m->n_boundary == 0)  // the reference panics inside merge_sort on an empty boundary (mesh_builder.rs:623-632)
    rc = fail(DZB_MESH_PARSE, "the mesh has no boundary edge");
  cudaStreamSynchronize(g_stream);
  dfree(d_xy);
  if (rc) {
    dzb_mesh2d_destroy(m);
    return rc;
  }
  *out = m;
  return DZB_OK;
}
int dzb_mesh2d_counts(const dzb_mesh2d* m, size_t* n_vertices, size_t* n_triangles, size_t* n_boundary) {
  if (!m) return fail(DZB_INVALID, "null pointer");
  if (n_vertices) *n_vertices = m->n_vertices;
  if (n_triangles) *n_triangles = m->n_triangles;
  if (n_boundary) *n_boundary = m->n_boundary;
  return DZB_OK;
}
int dzb_mesh2d_vertices(const dzb_mesh2d* m, double* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (len != 6 * m->n_vertices) return fail(DZB_WRONG_DIMS, "vertex buffer length (6 V)");
  CK(cudaMemcpyAsync(out, m->d_vertices, len * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh2d_indices(const dzb_mesh2d* m, uint32_t* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (len != 3 * m->n_triangles) return fail(DZB_WRONG_DIMS, "index buffer length (3 T)");
  CK(cudaMemcpyAsync(out, m->d_indices, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh2d_boundary_indices(const dzb_mesh2d* m, uint32_t* out, size_t len) {
  if (!m || !out) return fail(DZB_INVALID, "null pointer");
  if (len != m->n_boundary) return fail(DZB_WRONG_DIMS, "boundary buffer length");
  CK(cudaMemcpyAsync(out, m->d_boundary, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_mesh2d_info(const dzb_mesh2d* m, double* max_length, float translation[3]) {
  if (!m) return fail(DZB_INVALID, "null pointer");
  if (max_length) *max_length = m->max_length;
  if (translation)
    for (int k = 0; k < 3; ++k) translation[k] = m->translation[k];
  return DZB_OK;
}
void dzb_mesh2d_destroy(dzb_mesh2d* m) {
  if (!m) return;
  dfree(m->d_vertices), dfree(m->d_indices), dfree(m->d_boundary);
  delete m;
}
// the detection alone, on a device-resident triangle list (large meshes that never were an .obj file)
int dzb_boundary_vertices_device(const uint32_t* d_triangles, size_t n_triangles, size_t n_vertices, uint32_t* d_out, size_t* n_out) {
  if (!d_triangles || !d_out || !n_out) return fail(DZB_INVALID, "null pointer");
  return boundary_vertices_device(d_triangles, n_triangles, n_vertices, d_out, n_out);
}

// ------------------------------------------------------------------------------------------------- static solvers
static int block_assemble(dzb_solver* s, const double* x_with_halo);
static int static_alloc(dzb_solver* s, size_t n) {
  int rc;
  for (double** p : {&s->lo, &s->di, &s->up, &s->rhs, &s->out})
    if ((rc = dalloc(p, n))) return rc;
  return DZB_OK;
}
// Per-matrix work after (re)assembly: pick the solver the matrix allows, then factor for the serial path.
static int static_factor(dzb_solver* s, unsigned long long version) {
  int rc;
  s->parallel = s->want_parallel;
  if (s->want_parallel) {
    bool ok = false;
    if (version && s->classified_valid && s->classified_version == version) {
      ok = s->classified_ok;  // re-assembly of unchanged coordinates: identical bands, identical answer
    } else {
      if ((rc = tri_parallel_ok(s->lo, s->di, s->up, s->n, &ok))) return rc;
      s->classified_valid = version != 0, s->classified_version = version, s->classified_ok = ok;
    }
    if (!ok) s->parallel = false;  // indefinite / convection-dominated system: the reference's serial recurrence
  }
  if (s->parallel) {
    if (s->big.n != s->n && dzb::bigsys_prepare(&s->big, s->n)) return fail(DZB_CUDA, dzb::bigsys_error());
  } else {
    if (!s->c && (rc = dalloc(&s->c, s->n))) return rc;
    if (!s->den && (rc = dalloc(&s->den, s->n))) return rc;
    if (!s->rden && (rc = dalloc(&s->rden, s->n))) return rc;
    dzb::k_thomas_factor_warp<<<1, 32, 0, g_stream>>>(s->lo, s->di, s->up, s->c, s->den, s->rden, (long long)s->n);
    LAUNCHED();
  }
  return DZB_OK;
}
static int static_finish(dzb_solver* s, const dzb_options& opt, const dzb_mesh* m) {
  const size_t n = s->n;
  int solve_mode = opt.solve_mode;
  if (solve_mode == DZB_SOLVE_AUTO) solve_mode = n <= 65536 ? DZB_SOLVE_FAITHFUL : DZB_SOLVE_PARALLEL;
  // The Stokes matrix has a cancelling (~0) diagonal: only the serial recurrence started from row 0 is stable on it.
  if (s->kind == dzb::KIND_STOKES) solve_mode = DZB_SOLVE_FAITHFUL;
  s->want_parallel = solve_mode == DZB_SOLVE_PARALLEL;
  s->refine_steps = opt.refine_steps >= 0 ? opt.refine_steps : 0;  // 1e-11 of the exact solution at 10^7 rows without any
  (void)n;
  int rc = static_factor(s, m ? m->version : 0);
  if (rc) return rc;
  return sync_stream();
}

// bands of the coordinates the handle currently stands for, and the per-matrix choice of solver that goes with them
static int static_materialise(dzb_solver* s) {
  const double* x = s->x_view ? s->x_view : (s->src_mesh ? s->src_mesh->d_nodes : nullptr);
  if (!x) return fail(DZB_INVALID, "the solver has lost its node coordinates");
  dzb::AsmParams P{s->mu, s->b, s->bc0, s->bc1, 0.0, 0.0};
  dzb::AsmOut o{s->lo, s->di, s->up, s->rhs, nullptr, nullptr, nullptr};
  int rc = launch_assembly<dzb::KIND_TI>(x, s->n, 1, s->G, s->assembly_mode, P, o, nullptr);
  if (rc) return rc;
  s->bands_stale = false;
  return static_factor(s, s->x_version);
}
// Assemble + solve in one launch straight from the node coordinates.  *done = false: the path does not apply to this
// system (matrix class, size, no cooperative launch) and the caller goes on with materialised bands.
static int static_solve_onelaunch(dzb_solver* s, bool* done, const dzb::PeerBox* px = nullptr) {
  *done = false;
  // `known`: the kernel itself has looked at these coordinates (every row covered by the closed form, matrix class).  A class
  // taken from materialised bands (at creation) says nothing about the former, so it only serves to rule the path OUT.
  const bool known = s->ol_checked_valid && s->ol_checked_version == s->x_version;
  if (!s->is_block) {
    if (s->classified_valid && s->classified_version == s->x_version && !s->classified_ok) return DZB_OK;
    if (s->fused_bad_valid && s->fused_bad_version == s->x_version) return DZB_OK;
  }
  dzb::OneLaunchArgs a{};
  int rc = get_rule(s->G, &a.rule);
  if (rc) return rc;
  if (a.rule.m <= 0) return DZB_OK;
  double* r = s->ol_red;
  constexpr size_t RS = 8 * dzb::OL_P1M;
  a.k = dzb::TileArgs{s->x_view, nullptr, nullptr, nullptr, (long long)s->n, r, r + RS, r + 2 * RS, r + 3 * RS, nullptr, s->out,
                      s->is_block ? s->block_flags : 0};
  a.k.bc0 = s->bc0, a.k.bc1 = s->bc1;
  a.P = dzb::AsmParams{s->mu, s->b, s->bc0, s->bc1, 0.0, 0.0};
  a.n_local = (long long)(s->is_block ? s->n_alloc : s->n), a.halo_left = (s->is_block && (s->block_flags & dzb::BS_OPEN_LEFT)) ? 1 : 0;
  if (px) a.px = *px;
  a.s_ex = s->ol_s;
  {
    const size_t rb = dzb::onelaunch_rows_b(s->n), rc = dzb::onelaunch_rows_c(s->n);
    double* q = s->ol_pool;
    for (int i = 0; i < 4; ++i) a.gb[i] = q, q += rb;
    a.ub = q, q += rb;
    for (int i = 0; i < 4; ++i) a.gc[i] = q, q += rc;
    a.uc = q;
  }
  const bool trace_on = g_ol_trace;  // tuning aid: phase timeline of every CTA to stderr
  unsigned long long*& d_trace = device_ctx().d_trace;
  if (trace_on && !d_trace) CK(cudaMalloc((void**)&d_trace, 8 * 1024 * sizeof(unsigned long long)));
  a.trace = trace_on ? d_trace : nullptr;
  a.class_flags = known ? nullptr : s->d_cls;
  if (!known) CK(cudaMemsetAsync(s->d_cls, 0, 2 * sizeof(int), g_stream));
  int ol_grid = 0;
  const int lr = dzb::onelaunch_run(&s->big, a, s->ol_caps, g_stream, &g_launches, &ol_grid);
  if (lr > 0) return fail(DZB_CUDA, dzb::bigsys_error());
  if (lr < 0) {
    s->fused = false;
    return DZB_OK;
  }
  if (s->is_block) {  // a collective: never waits for the host; the flags are looked at by dzb_solver_block_read / _status
    if (!known) s->cls_pending = true, s->ol_checked_valid = true, s->ol_checked_version = s->x_version;
    *done = true;
    return DZB_OK;
  }
  if (trace_on) {
    std::vector<unsigned long long> h(8 * 1024);
    CK(cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    const int ctas = std::max(1, std::min(1024, ol_grid));
    unsigned long long t0 = ~0ULL;
    for (int c = 0; c < ctas; ++c) t0 = std::min(t0, h[8 * c]);
    if (const char* path = std::getenv("DZB_OL_RED_FILE")) {  // the middle system (2 rows per CTA), appended per solve
      constexpr size_t RS = 8 * dzb::OL_P1M;
      std::vector<double> red(4 * RS);
      CK(cudaMemcpy(red.data(), s->ol_red, red.size() * sizeof(double), cudaMemcpyDeviceToHost));
      if (FILE* f = std::fopen(path, "a")) {
        for (int r = 0; r < 2 * ctas; ++r) std::fprintf(f, "%.17g %.17g %.17g %.17g\n", red[r], red[RS + r], red[2 * RS + r], red[3 * RS + r]);
        std::fclose(f);
      }
    }
    if (const char* path = std::getenv("DZB_OL_TRACE_FILE")) {  // one line per CTA: the six stamps in microseconds
      if (FILE* f = std::fopen(path, "w")) {
        for (int c = 0; c < ctas; ++c) {  // the stamps in microseconds, then the SM the CTA ran on
          for (int k = 0; k < 8; ++k) std::fprintf(f, "%.3f ", (double)((k ? h[8 * c + k] : (h[8 * c] & ~0xfffULL)) - (t0 & ~0xfffULL)) * 1e-3);
          std::fprintf(f, "%u\n", (unsigned)(h[8 * c] & 0xfffULL));
        }
        std::fclose(f);
      }
    }
    const char* names[] = {"start", "pass1 end", "cta level reduced", "after grid sync", "middle solved", "pass2 end", "level B reduced", "level C reduced"};
    for (int k = 0; k < 8; ++k) {
      double lo = 1e30, hi = 0, sum = 0;
      for (int c = 0; c < ctas; ++c) {
        const double v = (double)(h[8 * c + k] - t0) * 1e-3;
        lo = std::min(lo, v), hi = std::max(hi, v), sum += v;
      }
      std::fprintf(stderr, "[ol trace] %-18s min %8.2f avg %8.2f max %8.2f us (%d CTAs)\n", names[k], lo, sum / ctas, hi, ctas);
    }
  }
  if (!known) {  // first solve of these coordinates: is the matrix one the partitioned elimination is reliable on?
    int f[2] = {0, 0};
    CK(cudaMemcpyAsync(f, s->d_cls, sizeof f, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (f[1]) {  // some row needs the quadrature loop: its values never existed on chip, nothing is known about the class
      s->fused_bad_valid = true, s->fused_bad_version = s->x_version;
      return DZB_OK;
    }
    s->classified_valid = true, s->classified_version = s->x_version, s->classified_ok = f[0] != 3;
    if (!s->classified_ok) return DZB_OK;
    s->ol_checked_valid = true, s->ol_checked_version = s->x_version;
  }
  *done = true;
  return DZB_OK;
}

int dzb_diffusion_ti_create(const dzb_mesh* m, double mu, double b, double bc_left, double bc_right, size_t G,
                            const dzb_options* opt_in, dzb_solver** out) {
  if (!out || !m) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  const dzb_options opt = resolve_options(opt_in);
  if (!options_valid(opt)) return fail(DZB_INVALID, "bad dzb_options");
  if (m->n < 3) return fail(DZB_WRONG_DIMS, "a bar needs at least 3 nodes");
  dzb_solver* s = new dzb_solver;
  s->kind = dzb::KIND_TI, s->n = m->n;
  s->mu = mu, s->b = b, s->bc0 = bc_left, s->bc1 = bc_right, s->G = G, s->assembly_mode = opt.assembly_mode;
  int rc = static_alloc(s, s->n);
  if (!rc) {
    dzb::AsmParams P{mu, b, bc_left, bc_right, 0.0, 0.0};
    dzb::AsmOut o{s->lo, s->di, s->up, s->rhs, nullptr, nullptr, nullptr};
    rc = launch_assembly<dzb::KIND_TI>(m->d_nodes, s->n, 1, G, opt.assembly_mode, P, o, nullptr);
  }
  if (!rc) rc = static_finish(s, opt, m);
  // closed-form assembly + partitioned solver on a bar of at least two tiles: from now on one launch per assemble + solve
  if (!rc && opt.assembly_mode == DZB_ASSEMBLY_FAST && s->want_parallel && s->refine_steps == 0 && s->n >= 2 * (size_t)dzb::BS_TILE &&
      !env_set("DZB_NO_ONELAUNCH")) {
    rc = dalloc(&s->ol_red, 4 * 8 * (size_t)dzb::OL_P1M);
    if (!rc) rc = dalloc(&s->ol_s, s->n);
    const size_t pool = 5 * (dzb::onelaunch_rows_b(s->n) + dzb::onelaunch_rows_c(s->n));
    if (!rc) rc = dalloc(&s->ol_pool, pool);
    // the f arrays of the levels B and C are only ever written where a Dirichlet row's image lands: zero everywhere else, for good
    if (!rc && cudaMemsetAsync(s->ol_pool, 0, pool * sizeof(double), g_stream) != cudaSuccess) rc = fail(DZB_CUDA, "cudaMemsetAsync");
    if (!rc) rc = dalloc(&s->d_cls, 2);
    if (!rc) {
      s->fused = true;
      solver_attach(s, m);
    }
  }
  if (!rc) s->x_version = m->version;
  if (rc) {
    dzb_solver_destroy(s);
    return rc;
  }
  *out = s;
  return DZB_OK;
}

int dzb_stokes1d_create(const dzb_mesh* m, double rho, double pressure, double (*force)(double, void*), void* ctx, size_t G,
                        const dzb_options* opt_in, dzb_solver** out) {
  if (!out || !m || !force) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  const dzb_options opt = resolve_options(opt_in);
  if (!options_valid(opt)) return fail(DZB_INVALID, "bad dzb_options");
  const size_t n = m->n;
  if (n < 3) return fail(DZB_WRONG_DIMS, "a bar needs at least 3 nodes");
  dzb::QuadRule rule;
  if (!dzb::gl_build_rule(G, &rule)) return fail(DZB_INTEGRATION, "Misuse of quad_pair function: k should be smaller than n");
  // Sample the closure on the host at the reference's points and in its order (dim1.rs:156-185 then :210-228):
  // T(x) = ((end-beg)/2)*x + (end+beg)/2, two roundings (host code is built with -ffp-contract=off).
  const size_t mq = rule.m;
  std::vector<double> F(mq * n, 0.0);
  const std::vector<double>* xp = nullptr;
  if (int hrc = host_nodes(m, &xp)) return hrc;
  const std::vector<double>& x = *xp;
  for (size_t i = 1; i + 1 < n; ++i) {
    const double cs = (x[i + 1] - x[i - 1]) / 2.0, ms = (x[i + 1] + x[i - 1]) / 2.0;
    for (size_t k = 0; k < mq; ++k) F[k * n + i] = force(cs * rule.x[k] + ms, ctx);
  }
  {
    const double cs = (x[1] - x[0]) / 2.0, ms = (x[1] + x[0]) / 2.0;
    for (size_t k = 0; k < mq; ++k) F[k * n] = force(cs * rule.x[k] + ms, ctx);
  }
  dzb_solver* s = new dzb_solver;
  s->kind = dzb::KIND_STOKES, s->n = n;
  double* d_force = nullptr;
  int rc = static_alloc(s, n);
  if (!rc) rc = dalloc(&d_force, F.size() + 1);
  if (!rc && !F.empty() &&
      cudaMemcpyAsync(d_force, F.data(), F.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
    rc = fail(DZB_CUDA, "H2D force samples");
  if (!rc) {
    dzb::AsmParams P{0.0, 0.0, 0.0, 0.0, rho, pressure};
    dzb::AsmOut o{s->lo, s->di, s->up, s->rhs, nullptr, nullptr, nullptr};
    rc = launch_assembly<dzb::KIND_STOKES>(m->d_nodes, n, 1, G, opt.assembly_mode, P, o, d_force);
  }
  if (!rc) rc = static_finish(s, opt, m);
  cudaStreamSynchronize(g_stream);
  dfree(d_force);
  if (rc) {
    dzb_solver_destroy(s);
    return rc;
  }
  *out = s;
  return DZB_OK;
}

int dzb_diffusion_td_create(const dzb_mesh* m, double mu, double b, double bc_left, double bc_right, const double* ic, size_t n_ic,
                            size_t G, const dzb_options* opt, dzb_solver** out) {
  if (!out || !m) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  dzb_ensemble* e = nullptr;
  const std::vector<double>* xp = nullptr;
  if (int hrc = host_nodes(m, &xp)) return hrc;
  int rc = ensemble_create_impl(xp->data(), 1, m->n, mu, b, bc_left, bc_right, ic, n_ic, G, opt, &e);
  if (rc) return rc;
  dzb_solver* s = new dzb_solver;
  s->kind = dzb::KIND_TD, s->n = m->n, s->td = e;
  *out = s;
  return DZB_OK;
}

int dzb_solver_rebuild(dzb_solver* s, const dzb_mesh* m) {
  if (!s || !m) return fail(DZB_INVALID, "null pointer");
  if (m->n != s->n) return fail(DZB_WRONG_DIMS, "rebuild needs a mesh with the solver's node count");
  if (s->is_block) return fail(DZB_INVALID, "block solvers are not rebuildable");
  if (s->kind == dzb::KIND_TI) {
    if (s->fused) {  // nothing to do yet: the rows are formed inside the solve, from these coordinates
      solver_attach(s, m);
      s->bands_stale = true;
      if (!(s->classified_valid && s->classified_version == m->version && !s->classified_ok)) return DZB_OK;
      return static_materialise(s);  // known not to allow the partitioned solver: bands + serial recurrence
    }
    dzb::AsmParams P{s->mu, s->b, s->bc0, s->bc1, 0.0, 0.0};
    dzb::AsmOut o{s->lo, s->di, s->up, s->rhs, nullptr, nullptr, nullptr};
    int rc = launch_assembly<dzb::KIND_TI>(m->d_nodes, s->n, 1, s->G, s->assembly_mode, P, o, nullptr);
    if (rc) return rc;
    s->x_version = m->version;
    return static_factor(s, m->version);
  }
  if (s->kind == dzb::KIND_TD) {
    dzb_ensemble* e = s->td;
    CK(cudaMemcpyAsync(e->d_mesh, m->d_nodes, s->n * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    return ensemble_assemble(e);
  }
  return fail(DZB_INVALID, "the Stokes solver samples a host closure: build a new solver instead");
}

int dzb_solver_solve_device(dzb_solver* s, double dt, const double** dev) {
  if (!s || !dev) return fail(DZB_INVALID, "null pointer");
  if (s->kind == dzb::KIND_TD) {
    int rc = ensemble_step_once(s->td, dt);
    if (rc) return rc;
    *dev = s->td->u[s->td->cur];
    return DZB_OK;
  }
  if (s->is_block) return fail(DZB_INVALID, "a block solver is driven through dzb_solver_block_* (it needs its neighbours)");
  if (s->fused && s->x_view) {
    bool done = false;
    int rc = static_solve_onelaunch(s, &done);
    if (rc) return rc;
    if (done) {
      s->last_path = 1;
      *dev = s->out;
      return DZB_OK;
    }
  }
  if (s->bands_stale) {  // the one-launch path did not apply: bands for the current coordinates, then the ordinary solvers
    int rc = static_materialise(s);
    if (rc) return rc;
  }
  s->last_path = 0;
  if (s->parallel) {
    if (dzb::bigsys_solve(&s->big, s->lo, s->di, s->up, s->rhs, s->out, s->refine_steps, g_stream, &g_launches))
      return fail(DZB_CUDA, dzb::bigsys_error());
  } else {
    dzb::k_thomas_solve_faithful_warp<<<1, 32, 0, g_stream>>>(s->lo, s->c, s->den, s->rden, s->rhs, s->out, (long long)s->n);
    LAUNCHED();
  }
  *dev = s->out;
  return DZB_OK;
}
int dzb_solver_solve(dzb_solver* s, double dt, double* out_host, size_t n) {
  if (!s || !out_host) return fail(DZB_INVALID, "null pointer");
  if (n != s->n) return fail(DZB_WRONG_DIMS, "output length != node count");
  const double* dev = nullptr;
  int rc = dzb_solver_solve_device(s, dt, &dev);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out_host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_solver_path(const dzb_solver* s, char* buf, size_t cap) {
  if (!s || !buf || cap == 0) return fail(DZB_INVALID, "null pointer");
  const char* name = "k_thomas_solve_faithful_warp (serial recurrence)";
  if (s->kind == dzb::KIND_TD) name = "time dependent: see dzb_ensemble_step_kernel";
  else if (s->last_path == 1) name = "k_tri_onelaunch<NODES> (assembly fused into the partitioned solve)";
  else if (s->parallel) name = "k_assemble_* + k_tri_* (materialised bands, partitioned solve)";
  std::snprintf(buf, cap, "%s", name);
  return DZB_OK;
}
int dzb_solver_node_count(const dzb_solver* s, size_t* n) {
  if (!s || !n) return fail(DZB_INVALID, "null pointer");
  *n = s->n;
  return DZB_OK;
}
int dzb_solver_bands(const dzb_solver* s, int which, double* lo, double* di, double* up, double* rhs, size_t n) {
  if (!s || !lo || !di || !up) return fail(DZB_INVALID, "null pointer");
  if (n != s->n) return fail(DZB_WRONG_DIMS, "band length != node count");
  const double *dl, *dd, *du, *dr = nullptr;
  if (s->kind == dzb::KIND_TD) {
    if (which == 1) dl = s->td->mlo, dd = s->td->mdi, du = s->td->mup;
    else if (which == 0) dl = s->td->slo, dd = s->td->sdi, du = s->td->sup;
    else return fail(DZB_INVALID, "which must be 0 (stiffness) or 1 (mass)");
  } else {
    if (which != 0) return fail(DZB_INVALID, "static solvers only have a stiffness matrix");
    if (s->bands_stale) {  // one-launch path: the bands only exist when somebody asks for them
      dzb_solver* ms = const_cast<dzb_solver*>(s);
      int rc = s->is_block ? block_assemble(ms, s->x_own + s->x_pad) : static_materialise(ms);
      if (rc) return rc;
    }
    dl = s->lo, dd = s->di, du = s->up, dr = s->rhs;
  }
  const size_t bytes = n * sizeof(double);
  CK(cudaMemcpyAsync(lo, dl, bytes, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(di, dd, bytes, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(up, du, bytes, cudaMemcpyDeviceToHost, g_stream));
  if (rhs && dr) CK(cudaMemcpyAsync(rhs, dr, bytes, cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_solver_state(const dzb_solver* s, double* out, size_t n) {
  if (!s || !out) return fail(DZB_INVALID, "null pointer");
  if (s->kind != dzb::KIND_TD) return fail(DZB_INVALID, "only the time-dependent solver has a state");
  if (n != s->n) return fail(DZB_WRONG_DIMS, "state length != node count");
  CK(cudaMemcpyAsync(out, s->td->u[s->td->cur], n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
void dzb_solver_destroy(dzb_solver* s) {
  if (!s) return;
  for (double** p : {&s->lo, &s->di, &s->up, &s->rhs}) {
    if (*p) *p -= s->off;
    dfree(*p);
  }
  for (double** p : {&s->c, &s->den, &s->rden, &s->out, &s->d_top8, &s->x_own, &s->ol_red, &s->ol_s, &s->ol_pool}) dfree(*p);
  solver_detach(s);
  if (s->d_cls) cudaFree(s->d_cls), s->d_cls = nullptr;
  dzb::bigsys_free(&s->big);
  ensemble_free(s->td);
  delete s;
}

// ------------------------------------------------------------------------------------------------- SPIKE blocks
int dzb_diffusion_ti_create_block(const dzb_mesh* m, int has_left, int has_right, double mu, double b, double bc_left,
                                  double bc_right, size_t G, const dzb_options* opt_in, dzb_solver** out) {
  if (!out || !m) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  const dzb_options opt = resolve_options(opt_in);
  if (!options_valid(opt)) return fail(DZB_INVALID, "bad dzb_options");
  const size_t halo = (has_left ? 1 : 0) + (has_right ? 1 : 0);
  if (m->n < 3 || m->n < halo + 2) return fail(DZB_WRONG_DIMS, "a block needs at least 2 own nodes besides its halo nodes");
  dzb_solver* s = new dzb_solver;
  s->kind = dzb::KIND_TI, s->n_alloc = m->n, s->n = m->n - halo;
  s->mu = mu, s->b = b, s->bc0 = bc_left, s->bc1 = bc_right, s->G = G, s->assembly_mode = opt.assembly_mode;
  s->is_block = true, s->parallel = true;
  s->block_flags = (has_left ? dzb::BS_OPEN_LEFT : 0) | (has_right ? dzb::BS_OPEN_RIGHT : 0);
  s->refine_steps = opt.refine_steps >= 0 ? opt.refine_steps : 0;
  // assemble over own + halo nodes: the halo rows come out as (discarded) boundary rows, every own row is exact
  // one spare element in front so that the block's own rows start 16-byte aligned (128-bit loads in k_tri_tile)
  const size_t lead = has_left ? 1 : 0;
  int rc = static_alloc(s, s->n_alloc + lead);
  if (!rc) rc = dalloc(&s->d_top8, 8);
  if (!rc) {
    dzb::AsmParams P{mu, b, bc_left, bc_right, 0.0, 0.0};
    dzb::AsmOut o{s->lo + lead, s->di + lead, s->up + lead, s->rhs + lead, nullptr, nullptr, nullptr};
    rc = launch_assembly<dzb::KIND_TI>(m->d_nodes, s->n_alloc, 1, G, opt.assembly_mode, P, o, nullptr);
  }
  if (!rc) {
    s->off = 2 * lead;
    s->lo += s->off, s->di += s->off, s->up += s->off, s->rhs += s->off;
    if (dzb::bigsys_prepare(&s->big, s->n)) rc = fail(DZB_CUDA, dzb::bigsys_error());
  }
  if (!rc) {
    bool ok = false;
    rc = tri_parallel_ok(s->lo, s->di, s->up, s->n, &ok, s->block_flags);
    if (!rc && !ok)
      rc = fail(DZB_INVALID, "the block's matrix is neither an M-matrix nor strictly dominant: the partitioned (SPIKE) solver "
                             "is not reliable on it; solve the bar on one GPU (serial recurrence)");
  }
  if (!rc) s->classified_valid = true, s->classified_version = m->version, s->classified_ok = true, s->x_version = m->version;
  // closed-form assembly on a block of at least two tiles: the rows can be formed inside the solve (k_tri_onelaunch), the
  // interface rows crossing GPUs through the peer mailboxes of a dzb_comm (dzb_solver_block_solve_device)
  if (!rc && opt.assembly_mode == DZB_ASSEMBLY_FAST && s->refine_steps == 0 && s->n >= 2 * (size_t)dzb::BS_TILE && !env_set("DZB_NO_ONELAUNCH")) {
    s->x_pad = lead;  // halo at [x_pad], first own row at [x_pad + lead]: index 0 or 2, 16-byte aligned
    rc = dalloc(&s->x_own, s->n_alloc + s->x_pad + 2);
    if (!rc) rc = dalloc(&s->ol_red, 4 * 8 * (size_t)dzb::OL_P1M);
    if (!rc) rc = dalloc(&s->ol_s, s->n);
    const size_t pool = 5 * (dzb::onelaunch_rows_b(s->n) + dzb::onelaunch_rows_c(s->n));
    if (!rc) rc = dalloc(&s->ol_pool, pool);
    // the f arrays of the levels B and C are only ever written where a Dirichlet row's image lands: zero everywhere else, for good
    if (!rc && cudaMemsetAsync(s->ol_pool, 0, pool * sizeof(double), g_stream) != cudaSuccess) rc = fail(DZB_CUDA, "cudaMemsetAsync");
    if (!rc) rc = dalloc(&s->d_cls, 2);
    if (!rc && cudaMemcpyAsync(s->x_own + s->x_pad, m->d_nodes, s->n_alloc * sizeof(double), cudaMemcpyDeviceToDevice, g_stream) != cudaSuccess)
      rc = fail(DZB_CUDA, "D2D node coordinates");
    if (!rc) s->x_view = s->x_own + s->x_pad + lead, s->fused = true;
  }
  if (!rc) rc = sync_stream();
  if (rc) {
    dzb_solver_destroy(s);
    return rc;
  }
  *out = s;
  return DZB_OK;
}
// bands of a block for its current coordinates (the halo rows come out as discarded boundary rows)
static int block_assemble(dzb_solver* s, const double* x_with_halo) {
  const size_t lead = (s->block_flags & dzb::BS_OPEN_LEFT) ? 1 : 0;
  dzb::AsmParams P{s->mu, s->b, s->bc0, s->bc1, 0.0, 0.0};
  dzb::AsmOut o{s->lo - lead, s->di - lead, s->up - lead, s->rhs - lead, nullptr, nullptr, nullptr};
  int rc = launch_assembly<dzb::KIND_TI>(x_with_halo, s->n_alloc, 1, s->G, s->assembly_mode, P, o, nullptr);
  if (!rc) s->bands_stale = false;
  return rc;
}
/* XSolver::new again for a block, into the buffers it owns: `m` holds the block's nodes + halo nodes as at creation. */
int dzb_solver_block_rebuild(dzb_solver* s, const dzb_mesh* m) {
  if (!s || !m) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  if (m->n != s->n_alloc) return fail(DZB_WRONG_DIMS, "rebuild needs a mesh with the block's node count (own + halo nodes)");
  if (s->fused) {  // the rows are formed inside the solve: only the (aligned, private) coordinates change
    if (m->version != s->x_version)
      CK(cudaMemcpyAsync(s->x_own + s->x_pad, m->d_nodes, s->n_alloc * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    s->x_version = m->version;
    s->bands_stale = true;
    return DZB_OK;
  }
  int rc = block_assemble(s, m->d_nodes);
  if (rc) return rc;
  s->x_version = m->version;
  if (!(s->classified_valid && s->classified_version == m->version)) {  // new coordinates: is the block still one the
    bool ok = false;                                                     // partitioned elimination is reliable on?
    if ((rc = tri_parallel_ok(s->lo, s->di, s->up, s->n, &ok, s->block_flags))) return rc;
    s->classified_valid = true, s->classified_version = m->version, s->classified_ok = ok;
    if (!ok) return fail(DZB_INVALID, "the block's matrix is neither an M-matrix nor strictly dominant");
  }
  return DZB_OK;
}
static int block_fresh_bands(dzb_solver* s) {  // one-launch blocks materialise their bands only when a caller needs them
  if (!s->bands_stale) return DZB_OK;
  return block_assemble(s, s->x_own + s->x_pad);
}
int dzb_solver_block_reduce(dzb_solver* s, int which_rhs, double top8[8]) {
  if (!s || !top8) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  if (int frc = block_fresh_bands(s)) return frc;
  const double* f = which_rhs ? s->big.resid : s->rhs;
  if (dzb::bigsys_block_reduce(&s->big, s->lo, s->di, s->up, f, s->block_flags, s->d_top8, g_stream, &g_launches))
    return fail(DZB_CUDA, dzb::bigsys_error());
  CK(cudaMemcpyAsync(top8, s->d_top8, 8 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_solver_block_backsolve(dzb_solver* s, int which_rhs, double us, double ue, int accumulate) {
  if (!s) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  const double* f = which_rhs ? s->big.resid : s->rhs;
  if (dzb::bigsys_block_backsolve(&s->big, s->lo, s->di, s->up, f, s->block_flags, us, ue, nullptr, s->out, accumulate != 0,
                                  g_stream, &g_launches))
    return fail(DZB_CUDA, dzb::bigsys_error());
  return DZB_OK;
}
int dzb_solver_block_edges(const dzb_solver* s, double edges[2]) {
  if (!s || !edges) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  CK(cudaMemcpyAsync(edges, s->out, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(edges + 1, s->out + s->n - 1, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_solver_block_residual(dzb_solver* s, double u_left, double u_right) {
  if (!s) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  dzb::k_residual_dd<<<dzb::bigsys_grid(s->n), 256, 0, g_stream>>>(s->lo, s->di, s->up, s->rhs, s->out, s->big.resid,
                                                                    (long long)s->n, s->block_flags, u_left, u_right, nullptr, nullptr);
  LAUNCHED();
  return DZB_OK;
}
// ---- the same five steps with every operand in DEVICE memory: nothing synchronises with the host, so reduce ->
// all-gather (NCCL, on the caller's stream) -> interface solve -> back-substitution queue up behind one another.
int dzb_solver_block_reduce_device(dzb_solver* s, int which_rhs, double* d_top8) {
  if (!s || !d_top8) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  if (int frc = block_fresh_bands(s)) return frc;
  const double* f = which_rhs ? s->big.resid : s->rhs;
  if (dzb::bigsys_block_reduce(&s->big, s->lo, s->di, s->up, f, s->block_flags, d_top8, g_stream, &g_launches))
    return fail(DZB_CUDA, dzb::bigsys_error());
  return DZB_OK;
}
int dzb_interface_solve_device(const double* d_rows, size_t n_blocks, double* d_ends, double* d_work) {
  if (!d_rows || !d_ends || !d_work) return fail(DZB_INVALID, "null pointer");
  if (n_blocks == 0 || n_blocks > 4096) return fail(DZB_WRONG_DIMS, "number of blocks");
  dzb::k_interface_solve<<<1, 32, 0, g_stream>>>(d_rows, (int)(2 * n_blocks), d_ends, d_work);
  LAUNCHED();
  return DZB_OK;
}
int dzb_solver_block_backsolve_device(dzb_solver* s, int which_rhs, const double* d_ends2, int accumulate) {
  if (!s || !d_ends2) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  const double* f = which_rhs ? s->big.resid : s->rhs;
  if (dzb::bigsys_block_backsolve(&s->big, s->lo, s->di, s->up, f, s->block_flags, 0.0, 0.0, d_ends2, s->out, accumulate != 0,
                                  g_stream, &g_launches))
    return fail(DZB_CUDA, dzb::bigsys_error());
  return DZB_OK;
}
int dzb_solver_block_edges_device(const dzb_solver* s, double* d_edges2) {
  if (!s || !d_edges2) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  CK(cudaMemcpyAsync(d_edges2, s->out, sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  CK(cudaMemcpyAsync(d_edges2 + 1, s->out + s->n - 1, sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  return DZB_OK;
}
int dzb_solver_block_residual_device(dzb_solver* s, const double* d_left, const double* d_right) {
  if (!s) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  dzb::k_residual_dd<<<dzb::bigsys_grid(s->n), 256, 0, g_stream>>>(s->lo, s->di, s->up, s->rhs, s->out, s->big.resid,
                                                                    (long long)s->n, s->block_flags, 0.0, 0.0, d_left, d_right);
  LAUNCHED();
  return DZB_OK;
}
int dzb_solver_block_device_solution(const dzb_solver* s, const double** device_ptr) {
  if (!s || !device_ptr) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  *device_ptr = s->out;
  return DZB_OK;
}
int dzb_solver_block_read(const dzb_solver* s, double* out_host, size_t n) {
  if (!s || !out_host) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  if (n != s->n) return fail(DZB_WRONG_DIMS, "output length != rows owned by the block");
  CK(cudaMemcpyAsync(out_host, s->out, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
// Host-side interface system of the SPIKE split: rows[8 r .. 8 r + 8) = block r's (a,c,s,f) s-row then e-row; unknowns
// in the order (s_0, e_0, s_1, e_1, ...) form a tridiagonal system of 2 n_blocks rows, solved here in long double.
int dzb_interface_solve(const double* rows, size_t n_blocks, double* ends) {
  if (!rows || !ends) return fail(DZB_INVALID, "null pointer");
  if (n_blocks == 0) return fail(DZB_WRONG_DIMS, "no blocks");
  const size_t m = 2 * n_blocks;
  std::vector<long double> a(m), bb(m), c(m), f(m), cp(m), dp(m);
  for (size_t i = 0; i < m; ++i) {
    a[i] = rows[4 * i], c[i] = rows[4 * i + 1], f[i] = rows[4 * i + 3];
    bb[i] = (long double)rows[4 * i + 2] - a[i] - c[i];  // diagonal = s - a - c
  }
  a[0] = 0.0L, c[m - 1] = 0.0L;  // closed ends of the whole bar (already 0 when the end blocks were created as such)
  cp[0] = c[0] / bb[0], dp[0] = f[0] / bb[0];
  for (size_t i = 1; i < m; ++i) {
    const long double den = bb[i] - a[i] * cp[i - 1];
    cp[i] = c[i] / den, dp[i] = (f[i] - a[i] * dp[i - 1]) / den;
  }
  long double x = dp[m - 1];
  ends[m - 1] = (double)x;
  for (size_t i = m - 1; i-- > 0;) {
    x = dp[i] - cp[i] * x;
    ends[i] = (double)x;
  }
  return DZB_OK;
}

}  // extern "C"
#include "dzb_comm.inl"
extern "C" {

// ------------------------------------------------------------------------------------------------- ensembles
int dzb_ensemble_td_create(const double* meshes, size_t n_bars, size_t n, double mu, double b, double bc_left, double bc_right,
                           const double* ic, size_t n_ic, size_t G, const dzb_options* opt, dzb_ensemble** out) {
  return ensemble_create_impl(meshes, n_bars, n, mu, b, bc_left, bc_right, ic, n_ic, G, opt, out);
}
int dzb_ensemble_step(dzb_ensemble* e, double dt, size_t steps) {
  if (!e) return fail(DZB_INVALID, "null pointer");
  size_t s = 0;
  if (steps > 1) {  // the first step also (re)builds the per-dt operators and decides whether the lean path applies
    int rc = ensemble_step_once(e, dt);
    if (rc) return rc;
    s = 1;
    // nobody can observe the intermediate states: keep operator and state on chip for blocks of steps (8 rows per lane only)
    if (e->fast && e->lean_ok && !e->use_big && e->lean_L <= 8 && !g_no_multi_step) {
      while (s < steps) {
        const int blk = (int)std::min<size_t>(steps - s, 1024);
        DISPATCH_LEAN(e, launch_step_lean_multi, e, blk);
        if (rc) return rc;
        e->cur ^= 1;
        s += (size_t)blk;
      }
    }
  }
  for (; s < steps; ++s) {
    int rc = ensemble_step_once(e, dt);
    if (rc) return rc;
  }
  return DZB_OK;
}
// one solve(dt) on every bar, all states to the host: as f64 (out64) or cast to f32 on the device first (out32)
static int ensemble_solve_impl(dzb_ensemble* e, double dt, double* out64, float* out32, size_t len) {
  if (!e || (!out64 && !out32)) return fail(DZB_INVALID, "null pointer");
  if (len != e->B * e->n) return fail(DZB_WRONG_DIMS, "output length != n_bars * n");
  int rc = DZB_OK;
  if (out32 && !e->u32) CK(cudaMalloc((void**)&e->u32, len * sizeof(float)));
  auto cast_slab = [&](const double* state, size_t first, size_t count) {  // state[first, first + count) -> e->u32
    const unsigned grid = (unsigned)std::min<size_t>((count / 2 + 255) / 256 + 1, 148 * 8);
    k_cast_f32<<<grid, 256, 0, g_stream>>>(state + first, e->u32 + first, (long long)count);
    ++g_launches;
  };
  // Steady state of a large ensemble on the lean path: the bars are stepped in slabs and every slab's 1/8 of the state
  // starts its way to the host (copy stream) while the next slab is still being stepped, so the step hides behind the
  // PCIe transfer instead of preceding it.  Same kernel, same arithmetic; anything else takes the plain path.
  if (e->fast && e->lean_ok && !e->use_big && e->prepared && e->prepared_dt == dt && e->B >= 256 * SLABS && !g_no_slab_copy) {
    DeviceCtx& cx = device_ctx();
    if (!cx.copy_ready) {  // a failed attempt leaves copy_ready false and is retried from where it stopped
      if (!cx.copy_stream) CK(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
      for (cudaEvent_t& ev : cx.stepped)
        if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      if (!cx.copied) CK(cudaEventCreateWithFlags(&cx.copied, cudaEventDisableTiming));
      cx.copy_ready = true;
    }
    cudaStream_t copy_stream = cx.copy_stream;
    cudaEvent_t* stepped = cx.stepped;
    cudaEvent_t copied = cx.copied;
    const size_t per = (e->B + SLABS - 1) / SLABS;
    for (size_t k = 0, bar0 = 0; k < SLABS && bar0 < e->B; ++k, bar0 += per) {
      const size_t nb = std::min(per, e->B - bar0);
      DISPATCH_LEAN(e, launch_step_lean_slab, e, bar0, nb);
      if (rc) return rc;
      if (out32) cast_slab(e->u[e->cur ^ 1], bar0 * e->n, nb * e->n);
      CK(cudaEventRecord(stepped[k], g_stream));
      CK(cudaStreamWaitEvent(copy_stream, stepped[k], 0));
      if (out32)
        CK(cudaMemcpyAsync(out32 + bar0 * e->n, e->u32 + bar0 * e->n, nb * e->n * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
      else
        CK(cudaMemcpyAsync(out64 + bar0 * e->n, e->u[e->cur ^ 1] + bar0 * e->n, nb * e->n * sizeof(double), cudaMemcpyDeviceToHost,
                           copy_stream));
    }
    e->cur ^= 1;
    CK(cudaEventRecord(copied, copy_stream));
    CK(cudaStreamWaitEvent(g_stream, copied, 0));  // later work on the library's stream is ordered after the read-back
    CK(cudaStreamSynchronize(copy_stream));
    return DZB_OK;
  }
  rc = ensemble_step_once(e, dt);
  if (rc) return rc;
  if (out32) {
    cast_slab(e->u[e->cur], 0, len);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out32, e->u32, len * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
  } else {
    CK(cudaMemcpyAsync(out64, e->u[e->cur], len * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  }
  return sync_stream();
}
int dzb_ensemble_solve(dzb_ensemble* e, double dt, double* out_host, size_t len) {
  if (!out_host) return fail(DZB_INVALID, "null pointer");
  return ensemble_solve_impl(e, dt, out_host, nullptr, len);
}
int dzb_ensemble_solve_f32(dzb_ensemble* e, double dt, float* out_host, size_t len) {
  if (!out_host) return fail(DZB_INVALID, "null pointer");
  return ensemble_solve_impl(e, dt, nullptr, out_host, len);
}
int dzb_ensemble_state(const dzb_ensemble* e, double* out_host, size_t len) {
  if (!e || !out_host) return fail(DZB_INVALID, "null pointer");
  if (len != e->B * e->n) return fail(DZB_WRONG_DIMS, "output length != n_bars * n");
  CK(cudaMemcpyAsync(out_host, e->u[e->cur], len * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_ensemble_read_bars(const dzb_ensemble* e, const size_t* ids, size_t count, double* out_host) {
  if (!e || !ids || !out_host) return fail(DZB_INVALID, "null pointer");
  for (size_t k = 0; k < count; ++k) {
    if (ids[k] >= e->B) return fail(DZB_WRONG_DIMS, "bar id out of range");
    CK(cudaMemcpyAsync(out_host + k * e->n, e->u[e->cur] + ids[k] * e->n, e->n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  }
  return sync_stream();
}
int dzb_ensemble_state_device(const dzb_ensemble* e, const double** p) {
  if (!e || !p) return fail(DZB_INVALID, "null pointer");
  *p = e->u[e->cur];
  return DZB_OK;
}
int dzb_ensemble_bands(const dzb_ensemble* e, size_t bar, int which, double* lo, double* di, double* up, size_t n) {
  if (!e || !lo || !di || !up) return fail(DZB_INVALID, "null pointer");
  if (n != e->n || bar >= e->B) return fail(DZB_WRONG_DIMS, "band length / bar id");
  if (which != 0 && which != 1) return fail(DZB_INVALID, "which must be 0 (stiffness) or 1 (mass)");
  const size_t o = bar * n, bytes = n * sizeof(double);
  CK(cudaMemcpyAsync(lo, (which ? e->mlo : e->slo) + o, bytes, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(di, (which ? e->mdi : e->sdi) + o, bytes, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(up, (which ? e->mup : e->sup) + o, bytes, cudaMemcpyDeviceToHost, g_stream));
  return sync_stream();
}
int dzb_ensemble_dims(const dzb_ensemble* e, size_t* n_bars, size_t* n) {
  if (!e) return fail(DZB_INVALID, "null pointer");
  if (n_bars) *n_bars = e->B;
  if (n) *n = e->n;
  return DZB_OK;
}
int dzb_ensemble_step_kernel(const dzb_ensemble* e, char* buf, size_t cap) {
  if (!e || !buf || cap == 0) return fail(DZB_INVALID, "null pointer");
  char name[96];
  if (e->use_big) std::snprintf(name, sizeof name, "k_td_rhs + k_tri_* (one long bar)");
  else if (e->fast && e->lean_ok) std::snprintf(name, sizeof name, "k_td_step_lean<%d,%d>", e->lean_L, e->lean_wpb);
  else if (e->fast) std::snprintf(name, sizeof name, "k_td_step_fast<%d,4>", e->L);
  else std::snprintf(name, sizeof name, (long long)e->B <= tw_max_bars() ? "k_td_step_faithful_warp" : "k_td_step_faithful");
  std::snprintf(buf, cap, "%s", name);
  return DZB_OK;
}
void dzb_ensemble_destroy(dzb_ensemble* e) { ensemble_free(e); }

// ---- EulerSolver ensembles ---------------------------------------------------------------------------------------------
int dzb_euler_create(size_t n_traj, size_t width, const double* coeffs, const double* values, dzb_euler** out) {
  if (!out || !coeffs || !values) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  if (width < 2 || width > 5) return fail(DZB_INVALID, "EulerSolver takes [f64; 2] .. [f64; 5] (solvers/euler/mod.rs:19-22)");
  if (n_traj == 0) return fail(DZB_INVALID, "no trajectories");
  dzb_euler* e = new dzb_euler;
  e->T = n_traj, e->W = width;
  int rc;
  if ((rc = dalloc(&e->v, n_traj * width)) || (rc = dalloc(&e->c, n_traj * (width + 1)))) {
    dzb_euler_destroy(e);
    return rc;
  }
  // host AoS -> device SoA
  std::vector<double> hv(n_traj * width), hc(n_traj * (width + 1));
  for (size_t i = 0; i < n_traj; ++i) {
    for (size_t j = 0; j < width; ++j) hv[j * n_traj + i] = values[i * width + j];
    for (size_t j = 0; j <= width; ++j) hc[j * n_traj + i] = coeffs[i * (width + 1) + j];
  }
  if (cudaMemcpyAsync(e->v, hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
      cudaMemcpyAsync(e->c, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
      (rc = sync_stream())) {
    dzb_euler_destroy(e);
    return rc ? rc : fail(DZB_CUDA, "H2D euler state");
  }
  *out = e;
  return DZB_OK;
}
int dzb_euler_create_program(size_t n_traj, size_t width, const dzb_euler_op* program, size_t n_ops, const double* params,
                             size_t n_params, const double* values, dzb_euler** out) {
  if (!out || !program || !values || (n_params && !params)) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  if (width < 2 || width > 5) return fail(DZB_INVALID, "EulerSolver takes [f64; 2] .. [f64; 5] (solvers/euler/mod.rs:19-22)");
  if (n_traj == 0) return fail(DZB_INVALID, "no trajectories");
  if (n_ops == 0 || n_ops > (size_t)dzb::EULER_MAX_OPS) return fail(DZB_INVALID, "a derivative program has 1 .. 96 operations");
  // run the program's stack discipline once on the host: the device interpreter does not check anything
  int depth = 0;
  std::vector<dzb::EulerOp> ops(n_ops);
  for (size_t q = 0; q < n_ops; ++q) {
    const int oc = program[q].opcode, arg = program[q].arg;
    if (oc < 0 || oc >= dzb::EOP_COUNT) return fail(DZB_INVALID, "unknown opcode in the derivative program");
    if (oc == dzb::EOP_VALUE && (arg < 0 || (size_t)arg >= width)) return fail(DZB_INVALID, "the program reads a value outside [f64; width]");
    if (oc == dzb::EOP_PARAM && (arg < 0 || (size_t)arg >= n_params)) return fail(DZB_INVALID, "the program reads a parameter that was not given");
    if (oc <= dzb::EOP_PARAM) ++depth;
    else if (oc <= dzb::EOP_DIV) {
      if (depth < 2) return fail(DZB_INVALID, "the derivative program pops an empty stack");
      --depth;
    } else if (depth < 1) return fail(DZB_INVALID, "the derivative program pops an empty stack");
    if (depth > dzb::EULER_MAX_STACK) return fail(DZB_INVALID, "the derivative program needs more than 16 stack entries");
    ops[q] = dzb::EulerOp{oc, arg, program[q].value};
  }
  if (depth != 1) return fail(DZB_INVALID, "the derivative program must leave exactly one value");
  dzb_euler* e = new dzb_euler;
  e->T = n_traj, e->W = width, e->n_ops = (int)n_ops, e->n_params = (int)n_params;
  int rc;
  if ((rc = dalloc(&e->v, n_traj * width)) || (rc = dalloc(&e->d_ops, n_ops)) || (rc = dalloc(&e->d_params, n_traj * n_params))) {
    dzb_euler_destroy(e);
    return rc;
  }
  std::vector<double> hv(n_traj * width), hp(n_traj * n_params);
  for (size_t i = 0; i < n_traj; ++i) {
    for (size_t j = 0; j < width; ++j) hv[j * n_traj + i] = values[i * width + j];
    for (size_t j = 0; j < n_params; ++j) hp[j * n_traj + i] = params[i * n_params + j];
  }
  if (cudaMemcpyAsync(e->v, hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
      cudaMemcpyAsync(e->d_ops, ops.data(), ops.size() * sizeof(dzb::EulerOp), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
      (!hp.empty() && cudaMemcpyAsync(e->d_params, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream) != cudaSuccess) ||
      (rc = sync_stream())) {
    dzb_euler_destroy(e);
    return rc ? rc : fail(DZB_CUDA, "H2D euler program");
  }
  *out = e;
  return DZB_OK;
}
int dzb_euler_set_values(dzb_euler* e, const double* values, size_t len) {
  if (!e || !values) return fail(DZB_INVALID, "null pointer");
  if (len != e->T * e->W) return fail(DZB_WRONG_DIMS, "length != n_traj * width");
  std::vector<double> hv(len);
  for (size_t i = 0; i < e->T; ++i)
    for (size_t j = 0; j < e->W; ++j) hv[j * e->T + i] = values[i * e->W + j];
  CK(cudaMemcpyAsync(e->v, hv.data(), len * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  return sync_stream();
}
int dzb_euler_step(dzb_euler* e, double step, size_t steps) {
  if (!e) return fail(DZB_INVALID, "null pointer");
  if (steps == 0) return DZB_OK;
  const long long T = (long long)e->T, k = (long long)steps;
  if (e->d_ops) {
    const dzb::EulerProgram prog{e->d_ops, e->n_ops, e->n_params, e->d_params};
    const unsigned pb = (unsigned)((e->T + 127) / 128);
    switch (e->W) {
      case 2: dzb::k_euler_program<2><<<pb, 128, 0, g_stream>>>(e->v, prog, T, step, k); break;
      case 3: dzb::k_euler_program<3><<<pb, 128, 0, g_stream>>>(e->v, prog, T, step, k); break;
      case 4: dzb::k_euler_program<4><<<pb, 128, 0, g_stream>>>(e->v, prog, T, step, k); break;
      default: dzb::k_euler_program<5><<<pb, 128, 0, g_stream>>>(e->v, prog, T, step, k); break;
    }
    LAUNCHED();
    return DZB_OK;
  }
  const unsigned blocks = (unsigned)((e->T + 255) / 256);
  switch (e->W) {
    case 2: dzb::k_euler_steps<2><<<blocks, 256, 0, g_stream>>>(e->v, e->c, T, step, k); break;
    case 3: dzb::k_euler_steps<3><<<blocks, 256, 0, g_stream>>>(e->v, e->c, T, step, k); break;
    case 4: dzb::k_euler_steps<4><<<blocks, 256, 0, g_stream>>>(e->v, e->c, T, step, k); break;
    default: dzb::k_euler_steps<5><<<blocks, 256, 0, g_stream>>>(e->v, e->c, T, step, k); break;
  }
  LAUNCHED();
  return DZB_OK;
}
int dzb_euler_values(const dzb_euler* e, double* out_host, size_t len) {
  if (!e || !out_host) return fail(DZB_INVALID, "null pointer");
  if (len != e->T * e->W) return fail(DZB_WRONG_DIMS, "output length != n_traj * width");
  std::vector<double> hv(len);
  CK(cudaMemcpyAsync(hv.data(), e->v, len * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  int rc = sync_stream();
  if (rc) return rc;
  for (size_t i = 0; i < e->T; ++i)
    for (size_t j = 0; j < e->W; ++j) out_host[i * e->W + j] = hv[j * e->T + i];
  return DZB_OK;
}
int dzb_euler_values_device(const dzb_euler* e, const double** device_ptr) {
  if (!e || !device_ptr) return fail(DZB_INVALID, "null pointer");
  *device_ptr = e->v;
  return DZB_OK;
}
void dzb_euler_destroy(dzb_euler* e) {
  if (!e) return;
  dfree(e->v);
  dfree(e->c);
  dfree(e->d_ops);
  dfree(e->d_params);
  delete e;
}

}  // extern "C"
