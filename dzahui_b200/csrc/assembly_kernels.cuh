// assembly_kernels.cuh — element integrals of the 1D linear-Lagrange FEM of Dzahui, on the device.
//
// Two evaluations of the same (quirky) reference quadrature (see DESIGN.md "assembly"):
//   FAITHFUL  the reference's loops, same operation order, separate mul/add (file compiled with -fmad=false):
//             diffusion_solver/time_independent.rs:102-171, time_dependent.rs:121-234, stokes_solver/dim1.rs:125-236.
//             One thread per matrix row, O(G) per row, bands bit-identical to the reference arithmetic.
//   FAST      closed form of those sums over prefix moments of the rule, keeping the dropped last node and the
//             floating-point split predicate `cs*x_k + ms < x_i` at the kink (piecewise_polynomials_1degree.rs:133-148).
//             O(log G) per row; HBM-bound: reads x (8 B/row), writes 4 (TI) or 6 (TD) bands.
// Layout: nodes and bands are bar-major [n_bars][n] f64 (n_bars = 1 for a single bar).
#pragma once
#include <cuda_runtime.h>

namespace dzb {

struct RuleView {      // device pointers into one allocation, see gl_quadrature.h for the meaning
  const double* x;     // [m]
  const double* w;     // [m]
  const double* u0;    // [m+1] ...
  const double* u1;
  const double* u2;
  const double* v0;
  const double* v1;
  const double* v2;
  int m;               // G - 1
  double t0, t1p, t1m, q;
  const unsigned short* guess;  // [GUESS_BINS + 1]: guess[k] = #{j : x_j >= -1 + 2 k / GUESS_BINS}, a bracket for the kink split
  // static diffusion, lean closed form (ti_elem / ti_row below)
  const double4* uv;             // [m + 1]: (u0, u1, v0, v1)[s], one 32-byte entry per split position
  const unsigned short* guessf;  // [GUESS_BINS + 2]: guess[k] with bit 15 set when a node lies within GUESS_TOL of bin k - 1
  double kv;                     // element validity factor: h * kv > max(|xa|, |xb|) puts every node image inside the element
};
constexpr int GUESS_BINS = 1024;
constexpr double GUESS_TOL = 0x1p-16;  // what separates a node from the kink for the table alone to decide the split

// ---- the reference's basis pieces, restated -----------------------------------------------------------------
// FirstDegreePolynomial (polynomials_1d.rs:16-19): evaluate = coefficient*x + independent_term (two roundings).
struct Affine {
  double c, t;
  __device__ __forceinline__ double operator()(double x) const { return c * x + t; }
};
__device__ __forceinline__ Affine to_0_1(double beg, double end) {  // transformation_to_0_1 :64-71
  return {1.0 / (end - beg), -beg / (end - beg)};
}
__device__ __forceinline__ Affine from_m1_p1(double beg, double end) {  // transformation_from_m1_p1 :74-81
  return {(end - beg) / 2.0, (end + beg) / 2.0};
}
__device__ __forceinline__ Affine compose(Affine s, Affine o) {  // self(other(x)) :116-121
  return {s.c * o.c, s.c * o.t + s.t};
}
// PiecewiseFirstDegreePolynomial (piecewise_polynomials_1degree.rs:17-20,133-148) of a hat function:
// np pieces (3 at the two ends, 4 inside), first and last identically zero.
struct Hat {
  Affine p[4];
  double bp[3];
  int np;
  __device__ __forceinline__ double eval(double x) const {
    const int nb = np - 1;
    for (int i = 0; i < nb; ++i)
      if (x < bp[i]) return p[i](x);
    return p[nb](x);
  }
  __device__ __forceinline__ double deriv(double x) const {  // differentiate(): pieces (0, c), same breakpoints
    const int nb = np - 1;
    for (int i = 0; i < nb; ++i)
      if (x < bp[i]) return 0.0 * x + p[i].c;
    return 0.0 * x + p[nb].c;
  }
};
// LinearBasis::new (linear_basis.rs:31-94), function i of a bar whose nodes are mesh[0..n)
__device__ __forceinline__ Hat make_hat(const double* __restrict__ mesh, long long n, long long i) {
  Hat h;
  const Affine zero = {0.0, 0.0}, phi1 = {1.0, 0.0}, phi2 = {-1.0, 1.0};
  if (i == 0) {
    h.np = 3;
    h.p[0] = zero, h.p[1] = compose(phi2, to_0_1(mesh[0], mesh[1])), h.p[2] = zero, h.p[3] = zero;
    h.bp[0] = mesh[0], h.bp[1] = mesh[1], h.bp[2] = 0.0;
  } else if (i == n - 1) {
    h.np = 3;
    h.p[0] = zero, h.p[1] = compose(phi1, to_0_1(mesh[n - 2], mesh[n - 1])), h.p[2] = zero, h.p[3] = zero;
    h.bp[0] = mesh[n - 2], h.bp[1] = mesh[n - 1], h.bp[2] = 0.0;
  } else {
    const double prev = mesh[i - 1], cur = mesh[i], next = mesh[i + 1];
    h.np = 4;
    h.p[0] = zero, h.p[1] = compose(phi1, to_0_1(prev, cur)), h.p[2] = compose(phi2, to_0_1(cur, next)), h.p[3] = zero;
    h.bp[0] = prev, h.bp[1] = cur, h.bp[2] = next;
  }
  return h;
}

enum { KIND_TI = 0, KIND_TD = 1, KIND_STOKES = 2 };

// 1/x by the hardware seed (MUFU.RCP64H, ~2^-23) and two Newton steps: no special cases and therefore no branch.
// Within an ulp of the correctly rounded value for normal x whose reciprocal is normal; garbage (inf / nan) otherwise.
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

struct AsmParams {
  double mu, b;          // diffusion
  double bc0, bc1;       // TI: Dirichlet values written to rhs
  double rho, pressure;  // Stokes
};
struct AsmOut {          // bar-major [n_bars][n]
  double *lo, *di, *up, *rhs;  // TI/Stokes: K and b_vector; TD: S (rhs unused)
  double *mlo, *mdi, *mup;     // TD: M
};

// ---- FAITHFUL row: the reference's inner loops for interior row i ---------------------------------------------
template <int KIND>
__device__ __forceinline__ void faithful_row(const double* __restrict__ mesh, long long n, long long i,
                                             const double* __restrict__ qx, const double* __restrict__ qw, int m,
                                             const AsmParams& P, const double* __restrict__ force, long long fstride,
                                             double out[7]) {
  const Hat phi = make_hat(mesh, n, i), phi_prev = make_hat(mesh, n, i - 1), phi_next = make_hat(mesh, n, i + 1);
  const Affine t_prev = from_m1_p1(mesh[i - 1], mesh[i]), t_next = from_m1_p1(mesh[i], mesh[i + 1]),
               t_square = from_m1_p1(mesh[i - 1], mesh[i + 1]);
  double a_prev = 0.0, a_next = 0.0, a_sq = 0.0;  // TI / Stokes integrals, or TD "prime"
  double h_prev = 0.0, h_next = 0.0, h_sq = 0.0;  // TD "half"
  double m_prev = 0.0, m_next = 0.0, m_sq = 0.0;  // TD "mass"
  double b_int = 0.0;                             // Stokes load
  for (int j = 0; j < m; ++j) {                   // `for j in 1..gauss_step`
    const double x = qx[j], w = qw[j];
    const double tp = t_prev(x), tn = t_next(x), ts = t_square(x);
    const double dtp = 0.0 * x + t_prev.c, dtn = 0.0 * x + t_next.c, dts = 0.0 * x + t_square.c;
    if (KIND == KIND_TI) {  // time_independent.rs:142-165
      a_prev += (P.mu * phi.deriv(tp) * phi_prev.deriv(tp) + P.b * phi_prev.deriv(tp) * phi.eval(tp)) * dtp * w;
      a_next += (P.mu * phi.deriv(tn) * phi_next.deriv(tn) + P.b * phi_next.deriv(tn) * phi.eval(tn)) * dtn * w;
      a_sq += (P.mu * phi.deriv(ts) * phi.deriv(ts) + P.b * phi.deriv(ts) * phi.eval(ts)) * dts * w;
    } else if (KIND == KIND_TD) {  // time_dependent.rs:180-218
      m_prev += phi.eval(tp) * phi_prev.eval(tp) * dtp * w;
      {
        const double v = phi.eval(ts);
        m_sq += (v * v) * dts * w;  // .powf(2.0)
      }
      m_next += phi.eval(tn) * phi_next.eval(tn) * dtn * w;
      a_prev += phi.deriv(tp) * phi_prev.deriv(tp) * dtp * w;
      {
        const double v = phi.deriv(ts);
        a_sq += (v * v) * dts * w;
      }
      a_next += phi.deriv(tn) * phi_next.deriv(tn) * dtn * w;
      h_prev += phi.eval(tp) * phi_prev.deriv(tp) * dtp * w;
      h_sq += phi.eval(ts) * phi.deriv(ts) * dts * w;
      h_next += phi.eval(tn) * phi_next.deriv(tn) * dtn * w;
    } else {  // dim1.rs:166-185
      a_prev += phi.eval(tp) * phi_prev.deriv(tp) * dtp * w;
      a_next += phi.eval(tn) * phi_next.deriv(tn) * dtn * w;
      a_sq += phi.eval(ts) * phi.deriv(ts) * dts * w;
      b_int += P.rho * force[(long long)j * fstride] * phi.eval(ts) * dts * w;
    }
  }
  if (KIND == KIND_TD) {
    out[0] = -P.mu * a_prev - P.b * h_prev;  // S   time_dependent.rs:226-233
    out[1] = -P.mu * a_sq - P.b * h_sq;
    out[2] = -P.mu * a_next - P.b * h_next;
    out[3] = 0.0;
    out[4] = m_prev, out[5] = m_sq, out[6] = m_next;  // M   :221-223
  } else {
    out[0] = a_prev, out[1] = a_sq, out[2] = a_next, out[3] = b_int;
    out[4] = out[5] = out[6] = 0.0;
  }
}

// ---- FAITHFUL row, lean: the same operations in the same order, with everything that does not depend on the
// quadrature node hoisted out of the loop.  Valid when every node image stays inside the element(s) it is meant for
// (checked with the reference's own two-rounding predicate at the two extreme nodes; fl(fl(cs*x)+ms) is monotone in x),
// so that the piecewise selects of `Hat::eval/deriv` are known per loop: then deriv() returns the piece's coefficient
// exactly (0.0*x + c == c) and eval() is one `c*x + t`.  Bands are bit-identical to faithful_row<KIND_TI>; rows that fail
// the check (degenerate or unsorted spacing) take the generic loop.  time_independent.rs:132-165.
__device__ __forceinline__ bool faithful_row_lean_ti(const double* __restrict__ mesh, long long n, long long i,
                                                     const double* __restrict__ qx, const double* __restrict__ qw, int m,
                                                     const AsmParams& P, double out[7]) {
  const double a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  const Affine tp = from_m1_p1(a, c), tn = from_m1_p1(c, e), ts = from_m1_p1(a, e);
  bool ok = m > 0 && (a < c) && (c < e);
  if (ok) {
    const double x_hi = qx[0], x_lo = qx[m - 1];
    ok = (tp(x_hi) < c) && !(tp(x_lo) < a) && (tn(x_hi) < e) && !(tn(x_lo) < c) && (ts(x_hi) < e) && !(ts(x_lo) < a);
    if (i >= 2) ok = ok && !(a < mesh[i - 2]);
    if (i + 2 < n) ok = ok && !(mesh[i + 2] < e);
  }
  if (!ok) return false;
  const Affine phi1 = {1.0, 0.0}, phi2 = {-1.0, 1.0};
  const Affine pl = compose(phi1, to_0_1(a, c));   // phi_i   on [a, c)
  const Affine pr = compose(phi2, to_0_1(c, e));   // phi_i   on [c, e)
  const Affine ql = compose(phi2, to_0_1(a, c));   // phi_i-1 on [a, c)
  const Affine qr = compose(phi1, to_0_1(c, e));   // phi_i+1 on [c, e)
  // (mu * phi_i' * phi_j') and (b * phi_j'), evaluated as the reference does for every node (same value every time)
  const double k1p = P.mu * pl.c * ql.c, k2p = P.b * ql.c;
  const double k1n = P.mu * pr.c * qr.c, k2n = P.b * qr.c;
  const double k1l = P.mu * pl.c * pl.c, k2l = P.b * pl.c;
  const double k1r = P.mu * pr.c * pr.c, k2r = P.b * pr.c;
  int lo = 0, hi = m;  // first node whose image under t_square is left of the kink (x_j decreases with j)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ts(qx[mid]) < c) hi = mid; else lo = mid + 1;
  }
  const int split = lo;
  double a_prev = 0.0, a_next = 0.0, a_sq = 0.0;
#pragma unroll 4
  for (int j = 0; j < split; ++j) {  // t_square(x_j) in [c, e): right piece
    const double x = qx[j], w = qw[j];
    a_prev += (k1p + k2p * pl(tp(x))) * tp.c * w;
    a_next += (k1n + k2n * pr(tn(x))) * tn.c * w;
    a_sq += (k1r + k2r * pr(ts(x))) * ts.c * w;
  }
#pragma unroll 4
  for (int j = split; j < m; ++j) {  // t_square(x_j) in [a, c): left piece
    const double x = qx[j], w = qw[j];
    a_prev += (k1p + k2p * pl(tp(x))) * tp.c * w;
    a_next += (k1n + k2n * pr(tn(x))) * tn.c * w;
    a_sq += (k1l + k2l * pl(ts(x))) * ts.c * w;
  }
  out[0] = a_prev, out[1] = a_sq, out[2] = a_next, out[3] = 0.0;
  out[4] = out[5] = out[6] = 0.0;
  return true;
}

// The time-dependent row the same way (time_dependent.rs:166-233): mass, "prime" and "half" integrals over the three
// maps, nine accumulators.  With the pieces known, phi'.phi'.dt is one node-independent number per loop, eval() is one
// affine evaluation: 46 fp64 operations per node instead of ~250.  Bit-identical to faithful_row<KIND_TD>.
__device__ __forceinline__ bool faithful_row_lean_td(const double* __restrict__ mesh, long long n, long long i,
                                                     const double* __restrict__ qx, const double* __restrict__ qw, int m,
                                                     const AsmParams& P, double out[7]) {
  const double a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  const Affine tp = from_m1_p1(a, c), tn = from_m1_p1(c, e), ts = from_m1_p1(a, e);
  bool ok = m > 0 && (a < c) && (c < e);
  if (ok) {
    const double x_hi = qx[0], x_lo = qx[m - 1];
    ok = (tp(x_hi) < c) && !(tp(x_lo) < a) && (tn(x_hi) < e) && !(tn(x_lo) < c) && (ts(x_hi) < e) && !(ts(x_lo) < a);
    if (i >= 2) ok = ok && !(a < mesh[i - 2]);
    if (i + 2 < n) ok = ok && !(mesh[i + 2] < e);
  }
  if (!ok) return false;
  const Affine phi1 = {1.0, 0.0}, phi2 = {-1.0, 1.0};
  const Affine pl = compose(phi1, to_0_1(a, c));   // phi_i   on [a, c)
  const Affine pr = compose(phi2, to_0_1(c, e));   // phi_i   on [c, e)
  const Affine ql = compose(phi2, to_0_1(a, c));   // phi_i-1 on [a, c)
  const Affine qr = compose(phi1, to_0_1(c, e));   // phi_i+1 on [c, e)
  // phi' * phi' * dt, formed as the loop forms it for every node (same operands, same roundings, every time)
  const double dd_p = pl.c * ql.c * tp.c, dd_n = pr.c * qr.c * tn.c;
  const double dd_l = (pl.c * pl.c) * ts.c, dd_r = (pr.c * pr.c) * ts.c;
  int lo = 0, hi = m;  // first node whose image under t_square is left of the kink (x_j decreases with j)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ts(qx[mid]) < c) hi = mid; else lo = mid + 1;
  }
  const int split = lo;
  double a_prev = 0.0, a_next = 0.0, a_sq = 0.0, h_prev = 0.0, h_next = 0.0, h_sq = 0.0, m_prev = 0.0, m_next = 0.0, m_sq = 0.0;
#pragma unroll 2
  for (int j = 0; j < split; ++j) {  // t_square(x_j) in [c, e): right piece
    const double x = qx[j], w = qw[j];
    const double ep = pl(tp(x)), eq = ql(tp(x)), en = pr(tn(x)), er = qr(tn(x)), v = pr(ts(x));
    m_prev += ep * eq * tp.c * w, a_prev += dd_p * w, h_prev += ep * ql.c * tp.c * w;
    m_next += en * er * tn.c * w, a_next += dd_n * w, h_next += en * qr.c * tn.c * w;
    m_sq += (v * v) * ts.c * w, a_sq += dd_r * w, h_sq += v * pr.c * ts.c * w;
  }
#pragma unroll 2
  for (int j = split; j < m; ++j) {  // t_square(x_j) in [a, c): left piece
    const double x = qx[j], w = qw[j];
    const double ep = pl(tp(x)), eq = ql(tp(x)), en = pr(tn(x)), er = qr(tn(x)), v = pl(ts(x));
    m_prev += ep * eq * tp.c * w, a_prev += dd_p * w, h_prev += ep * ql.c * tp.c * w;
    m_next += en * er * tn.c * w, a_next += dd_n * w, h_next += en * qr.c * tn.c * w;
    m_sq += (v * v) * ts.c * w, a_sq += dd_l * w, h_sq += v * pl.c * ts.c * w;
  }
  out[0] = -P.mu * a_prev - P.b * h_prev;  // S   time_dependent.rs:226-233
  out[1] = -P.mu * a_sq - P.b * h_sq;
  out[2] = -P.mu * a_next - P.b * h_next;
  out[3] = 0.0;
  out[4] = m_prev, out[5] = m_sq, out[6] = m_next;  // M   :221-223
  return true;
}

// Stokes row 0 (dim1.rs:195-236)
__device__ __forceinline__ void faithful_stokes_row0(const double* __restrict__ mesh, long long n,
                                                     const double* __restrict__ qx, const double* __restrict__ qw,
                                                     int m, const AsmParams& P, const double* __restrict__ force,
                                                     long long fstride, double out[7]) {
  const Hat phi0 = make_hat(mesh, n, 0), phi1 = make_hat(mesh, n, 1);
  const Affine t0 = from_m1_p1(mesh[0], mesh[1]);
  double i0 = 0.0, i0n = 0.0, bf = 0.0;
  for (int j = 0; j < m; ++j) {
    const double x = qx[j], w = qw[j];
    const double tr = t0(x), dt0 = 0.0 * x + t0.c;
    i0 += phi0.eval(tr) * phi0.deriv(tr) * dt0 * w;
    i0n += phi0.eval(tr) * phi1.deriv(tr) * dt0 * w;
    bf += P.rho * force[(long long)j * fstride] * phi0.eval(tr) * dt0 * w;
  }
  out[0] = 0.0, out[1] = i0, out[2] = i0n, out[3] = bf;
  out[4] = out[5] = out[6] = 0.0;
}

// Boundary rows shared by both modes.  TI: identity row + Dirichlet value (time_independent.rs:176-179);
// TD: identity rows in both M and S (time_dependent.rs:236-239); Stokes last row (dim1.rs:234,236).
template <int KIND>
__device__ __forceinline__ void boundary_row(bool first, const AsmParams& P, double out[7]) {
  out[0] = 0.0, out[1] = 1.0, out[2] = 0.0, out[3] = 0.0, out[4] = 0.0, out[5] = 0.0, out[6] = 0.0;
  if (KIND == KIND_TI) out[3] = first ? P.bc0 : P.bc1;
  if (KIND == KIND_TD) out[5] = 1.0;
  if (KIND == KIND_STOKES) out[3] = P.pressure;  // only the last row reaches here
}

template <int KIND>
__device__ __forceinline__ void store_row(const AsmOut& o, long long g, const double v[7]) {
  o.lo[g] = v[0], o.di[g] = v[1], o.up[g] = v[2];
  if (KIND != KIND_TD) o.rhs[g] = v[3];
  if (KIND == KIND_TD) o.mlo[g] = v[4], o.mdi[g] = v[5], o.mup[g] = v[6];
}

// row index -> (bar, row in bar) without a 64-bit division on the common paths
__device__ __forceinline__ void split_index(long long g, long long n, long long n_bars, long long* bar, long long* i) {
  if (n_bars == 1) {
    *bar = 0, *i = g;
  } else if ((g | n) >> 31 == 0) {
    const unsigned b = (unsigned)g / (unsigned)n;
    *bar = b, *i = (unsigned)g - b * (unsigned)n;
  } else {
    *bar = g / n, *i = g - *bar * n;
  }
}

// One thread per matrix row; the rule table is staged in shared memory when it fits (smem_rule != 0).
// Stokes: `force` holds rho-free samples f(T_square,i(x_k)) k-major: force[(k-1)*n + i], row 0 at i = 0.
template <int KIND>
__global__ void __launch_bounds__(128) k_assemble_faithful(const double* __restrict__ mesh, long long n,
                                                           long long n_bars, RuleView rule, AsmParams P, AsmOut o,
                                                           const double* __restrict__ force, int smem_rule) {
  extern __shared__ double s_rule[];
  const double* qx = rule.x;
  const double* qw = rule.w;
  if (smem_rule) {
    for (int j = threadIdx.x; j < rule.m; j += blockDim.x) s_rule[j] = rule.x[j], s_rule[rule.m + j] = rule.w[j];
    __syncthreads();
    qx = s_rule, qw = s_rule + rule.m;
  }
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * n_bars) return;
  long long bar, i;
  split_index(g, n, n_bars, &bar, &i);
  const double* bm = mesh + bar * n;
  double v[7];
  if (i == n - 1 || (i == 0 && KIND != KIND_STOKES))
    boundary_row<KIND>(i == 0, P, v);
  else if (i == 0)
    faithful_stokes_row0(bm, n, qx, qw, rule.m, P, force, n, v);
  else if (!(KIND == KIND_TI && faithful_row_lean_ti(bm, n, i, qx, qw, rule.m, P, v)) &&
           !(KIND == KIND_TD && faithful_row_lean_td(bm, n, i, qx, qw, rule.m, P, v)))
    faithful_row<KIND>(bm, n, i, qx, qw, rule.m, P, force ? force + i : nullptr, n, v);
  store_row<KIND>(o, g, v);
}

// ---- FAST row: closed form ----------------------------------------------------------------------------------
// In exact arithmetic the reference's integrands only depend on x_k through (1 +- x_k):
//   on [a,c] ("prev"):   phi_i = (1+x)/2, phi_{i-1} = (1-x)/2, phi_i' = 1/h1, phi_{i-1}' = -1/h1, Jacobian h1/2
//   on [c,e] ("next"):   phi_i = (1-x)/2, phi_{i+1} = (1+x)/2, phi_i' = -1/h2, phi_{i+1}' = 1/h2, Jacobian h2/2
//   on [a,e] ("square"): X_k < c -> phi_i = H(1+x)/(2 h1), phi_i' = 1/h1;  else phi_i = H(1-x)/(2 h2), phi_i' = -1/h2;
//                        Jacobian H/2, H = h1 + h2.  WHICH nodes are left of the kink is decided with the reference's
//                        own two-rounding predicate fl(fl(cs*x_k) + ms) < c, found by bisection (x_k decreasing in k).
// Falls back to the faithful loop for a row whose pieces could be selected differently by rounding (degenerate
// or unsorted spacing), so FAST never silently changes which polynomial piece a node hits.
// true when every node image of row i stays inside the element(s) it is meant for (so the closed forms apply)
__device__ __forceinline__ bool fast_row_ok(const double* __restrict__ mesh, long long n, long long i, double x_hi, double x_lo,
                                            int m) {
  const double a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  if (!(m > 0 && (a < c) && (c < e))) return false;
  const Affine tp = from_m1_p1(a, c), tn = from_m1_p1(c, e), ts = from_m1_p1(a, e);
  bool ok = (tp(x_hi) < c) && !(tp(x_lo) < a) && (tn(x_hi) < e) && !(tn(x_lo) < c) && (ts(x_hi) < e) && !(ts(x_lo) < a);
  if (i >= 2) ok = ok && !(a < mesh[i - 2]);
  if (i + 2 < n) ok = ok && !(mesh[i + 2] < e);
  return ok;
}

// Returns false (out untouched) for a row that must take the quadrature loop (k_assemble_fixup).
template <int KIND>
__device__ __forceinline__ bool fast_row(const double* __restrict__ mesh, long long n, long long i,
                                         const RuleView& R, const double* __restrict__ qx, const AsmParams& P, double out[7]) {
  const int m = R.m;
  if (m <= 0) return false;
  if (!fast_row_ok(mesh, n, i, qx[0], qx[m - 1], m)) return false;
  const double a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  const Affine ts = from_m1_p1(a, e);
  const double h1 = c - a, h2 = e - c, H = e - a;
  // First node index whose image is left of the kink.  The predicate fl(fl(cs x_j) + ms) < c is the reference's own and is
  // monotone in j (x_j decreases); a table over x* = (h1 - h2) / H brackets the answer, two probes settle it exactly.
  int s;
  {
    const float xs = __fdividef((float)(h1 - h2), (float)H);
    int bin = (int)((xs + 1.0f) * (0.5f * GUESS_BINS));
    bin = max(0, min(GUESS_BINS - 1, bin));
    s = R.guess[bin + 1];
    while (s < m && !(ts(qx[s]) < c)) ++s;
    while (s > 0 && (ts(qx[s - 1]) < c)) --s;
  }
  const double r1 = __drcp_rn(h1), r2 = __drcp_rn(h2);  // two reciprocals instead of four divisions
  const double ih1sq = r1 * r1, ih2sq = r2 * r2;
  const double U0 = R.u0[s], U1 = R.u1[s], V0 = R.v0[s], V1 = R.v1[s];
  const double prime_prev = -0.5 * R.t0 * r1, prime_next = -0.5 * R.t0 * r2;
  const double prime_sq = (0.5 * H) * fma(U0, ih1sq, V0 * ih2sq);
  const double half_prev = -0.25 * R.t1p, half_next = 0.25 * R.t1m;
  const double half_sq = (0.25 * H * H) * fma(U1, ih1sq, -V1 * ih2sq);
  if (KIND == KIND_TI) {
    out[0] = fma(P.mu, prime_prev, P.b * half_prev);
    out[1] = fma(P.mu, prime_sq, P.b * half_sq);
    out[2] = fma(P.mu, prime_next, P.b * half_next);
    out[3] = 0.0;
    out[4] = out[5] = out[6] = 0.0;
  } else {
    const double U2 = R.u2[s], V2 = R.v2[s];
    out[0] = -fma(P.mu, prime_prev, P.b * half_prev);
    out[1] = -fma(P.mu, prime_sq, P.b * half_sq);
    out[2] = -fma(P.mu, prime_next, P.b * half_next);
    out[3] = 0.0;
    out[4] = (0.125 * h1) * R.q;
    out[5] = (0.125 * H * H * H) * fma(U2, ih1sq, V2 * ih2sq);
    out[6] = (0.125 * h2) * R.q;
  }
  return true;
}

// ---- static diffusion, lean closed form ---------------------------------------------------------------------------------
// The same sums as fast_row<KIND_TI>, arranged per ELEMENT so that a thread owning consecutive rows (the fused assemble +
// solve kernel, onelaunch_kernels.cuh) forms every element quantity once, and with the two expensive parts of fast_row
// replaced by cheap sufficient tests:
//   * validity.  fast_row_ok evaluates the reference's own predicate at the two extreme nodes of all three maps (24 fp64
//     operations per row).  Here an element [xa, xb] passes when  h kv > M,  M = max(|xa|, |xb|),  2^-899 < h,  M < 2^1000,
//     with kv = min(delta 2^49, 2^30) and delta = min(1 - x_first, 1 + x_last) the distance of the rule's extreme nodes from
//     +-1.  Proof sketch: from_m1_p1 gives cm = h/2 exactly and mm = (xa + xb)(1 + e)/2, so the computed image
//     fl(fl(cm x) + mm) is within 6 u M of mid + half x (u = 2^-53, |x| <= 1); the exact images of the extreme nodes are
//     half delta inside the element, and h kv > M makes half delta > 6 u M (16 u M / delta < h).  The two-element map of
//     the diagonal integral has H >= h and M' <= max(M1, M2), so it passes when both elements do.  The range conditions keep
//     every intermediate normal.  Elements that fail are NOT necessarily bad: their rows simply take the quadrature loop.
//   * the kink split.  The table over x* = (h1 - h2) / H answers alone unless a node lies within GUESS_TOL of the bin
//     (bit 15 of guessf): h > 2^-30 M bounds the predicate's rounding to 1.4e-6 in x*, the float arithmetic adds 3e-7.
//     Flagged bins settle the split with the reference's own predicate, as fast_row does.
// Reciprocals: rcp_newton (branch-free).  Materialised bands (k_assemble_fast<KIND_TI>) and rows formed on chip use these
// very functions, so they agree bit for bit.
struct TiK {  // per (rule, parameters)
  double mu, b, K0, bhp, bhn, kv;
};
__device__ __forceinline__ TiK make_tik(const RuleView& R, const AsmParams& P) {
  return {P.mu, P.b, -0.5 * R.t0, P.b * (-0.25 * R.t1p), P.b * (0.25 * R.t1m), R.kv};
}
struct TiElem {
  double h, r, r2, k0r;  // length, 1/h, 1/h^2, K0/h
  bool ok;
};
__device__ __forceinline__ TiElem ti_elem(double xa, double xb, const TiK& K) {
  TiElem e;
  e.h = xb - xa;
  e.r = rcp_newton(e.h);
  e.r2 = e.r * e.r, e.k0r = K.K0 * e.r;
  const double M = fmax(fabs(xa), fabs(xb));
  e.ok = (e.h * K.kv > M) && (M < 0x1p1000) && (e.h > 0x1p-899);
  return e;
}
// flagged bin: the reference's predicate decides (kept out of line: the callers unroll eight rows)
__device__ __noinline__ int ti_split_probe(double xa, double xc, double xe, int s, const double* qx, int m) {
  const Affine ts = from_m1_p1(xa, xe);
  while (s < m && !(ts(qx[s]) < xc)) ++s;
  while (s > 0 && (ts(qx[s - 1]) < xc)) --s;
  return s;
}
// interior row between elements L = [xa, xc] and Rr = [xc, xe], both ok
__device__ __forceinline__ void ti_row(const TiElem& L, const TiElem& Rr, double xa, double xc, double xe, const TiK& K,
                                       const double4* __restrict__ uv, const unsigned short* __restrict__ guessf,
                                       const double* qx, int m, double& lo, double& di, double& up, bool live = true) {
  // live == false: the caller forms this row unconditionally (straight-line code) but will discard it: never probe for it
  const double H = xe - xa;
  const float xs = __fdividef((float)(L.h - Rr.h), (float)H);
  int bin = (int)((xs + 1.0f) * (0.5f * GUESS_BINS));
  bin = max(0, min(GUESS_BINS - 1, bin));
  int s = guessf[bin + 1];
  if ((s & 0x8000) && live) s = ti_split_probe(xa, xc, xe, s & 0x7fff, qx, m);
  s &= 0x7fff;
  const double4 t = uv[s];  // (U0, U1, V0, V1)
  const double prime_sq = (0.5 * H) * fma(t.x, L.r2, t.z * Rr.r2);
  const double half_sq = (0.25 * H * H) * fma(t.y, L.r2, -t.w * Rr.r2);
  lo = fma(K.mu, L.k0r, K.bhp);
  di = fma(K.mu, prime_sq, K.b * half_sq);
  up = fma(K.mu, Rr.k0r, K.bhn);
}
// whether interior row i takes the closed form: both elements pass, outer neighbours not out of order
__device__ __forceinline__ bool ti_row_ok(const double* __restrict__ mesh, long long n, long long i, const TiElem& L, const TiElem& Rr,
                                          double xa, double xe) {
  bool ok = L.ok && Rr.ok;
  if (i >= 2) ok = ok && !(xa < mesh[i - 2]);
  if (i + 2 < n) ok = ok && !(mesh[i + 2] < xe);
  return ok;
}
__device__ __forceinline__ bool fast_row_ti(const double* __restrict__ mesh, long long n, long long i, const RuleView& R,
                                            const double* qx, const TiK& K, double out[7]) {
  if (R.m <= 0) return false;
  const double xa = mesh[i - 1], xc = mesh[i], xe = mesh[i + 1];
  const TiElem L = ti_elem(xa, xc, K), Rr = ti_elem(xc, xe, K);
  if (!ti_row_ok(mesh, n, i, L, Rr, xa, xe)) return false;
  ti_row(L, Rr, xa, xc, xe, K, R.uv, R.guessf, qx, R.m, out[0], out[1], out[2]);
  out[3] = 0.0, out[4] = out[5] = out[6] = 0.0;
  return true;
}

// One thread per row, grid-stride; x table (and nothing else) staged in shared memory for the probes.  Rows the closed
// form does not cover are counted in *bad and left to k_assemble_fixup, which keeps the quadrature-loop code (and its
// local-memory frame) out of this kernel.
template <int KIND>
__global__ void __launch_bounds__(256) k_assemble_fast(const double* __restrict__ mesh, long long n, long long n_bars,
                                                       RuleView rule, AsmParams P, AsmOut o, int smem_rule,
                                                       unsigned* __restrict__ bad) {
  extern __shared__ double s_rule[];
  const double* qx = rule.x;
  if (smem_rule) {
    for (int j = threadIdx.x; j < rule.m; j += blockDim.x) s_rule[j] = rule.x[j];
    __syncthreads();
    qx = s_rule;
  }
  const long long total = n * n_bars;
  const TiK K = make_tik(rule, P);
  unsigned nbad = 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long bar, i;
    split_index(g, n, n_bars, &bar, &i);
    double v[7];
    if (i == 0 || i == n - 1)
      boundary_row<KIND>(i == 0, P, v);
    else if (!(KIND == KIND_TI ? fast_row_ti(mesh + bar * n, n, i, rule, qx, K, v) : fast_row<KIND>(mesh + bar * n, n, i, rule, qx, P, v))) {
      ++nbad;
      continue;
    }
    store_row<KIND>(o, g, v);
  }
  if (nbad) atomicAdd(bad, nbad);
}
// The rows k_assemble_fast skipped, through the reference's quadrature loop.  Launched unconditionally (no host
// synchronisation); returns at once when nothing was skipped.
template <int KIND>
__global__ void __launch_bounds__(128) k_assemble_fixup(const double* __restrict__ mesh, long long n, long long n_bars,
                                                        RuleView rule, AsmParams P, AsmOut o,
                                                        const unsigned* __restrict__ bad) {
  if (*bad == 0) return;
  const int m = rule.m;
  const TiK K = make_tik(rule, P);
  const long long total = n * n_bars;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long bar, i;
    split_index(g, n, n_bars, &bar, &i);
    if (i == 0 || i == n - 1) continue;
    const double* bm = mesh + bar * n;
    if (KIND == KIND_TI) {
      const double xa = bm[i - 1], xc = bm[i], xe = bm[i + 1];
      if (m > 0 && ti_row_ok(bm, n, i, ti_elem(xa, xc, K), ti_elem(xc, xe, K), xa, xe)) continue;
    } else if (m > 0 && fast_row_ok(bm, n, i, rule.x[0], rule.x[m - 1], m)) continue;
    double v[7];
    faithful_row<KIND>(bm, n, i, rule.x, rule.w, m, P, nullptr, 0, v);
    store_row<KIND>(o, g, v);
  }
}

}  // namespace dzb
