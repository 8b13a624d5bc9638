// dzahui_oracle.cpp — CPU ORACLE (test infrastructure, see dzahui_oracle.hpp header).
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  No -ffast-math, ever.
#include "dzahui_oracle.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "gl_tables_generated.h"

namespace dzo {

static const double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI

// ============================================================== quadrature
// Horner evaluation, highest-degree coefficient first.  `(a*x + b)*x - c` in the Rust source and
// `(a*x + b)*x + (-c)` here round identically (IEEE subtraction == addition of the negation).
static inline double horner(const double* c, int n, double x) {
  double acc = c[0];
  for (int i = 1; i < n; ++i) acc = acc * x + c[i];
  return acc;
}

// gauss_legendre.rs:5684-5702 — k-th zero of J0: table for k < 20, McMahon expansion otherwise.
double bessel_j_zero(size_t k) {
  if (k < 20) return dzb_gl_tables::JZ[k - 1];
  const double ak = PI * ((double)k - 0.25);
  const double r = 1.0 / ak;
  const double r2 = r * r;
  // ak + r*(0.125 + r2*(m1 + r2*(m2 + ... + r2*(m7 + m8*r2))))
  static const double m[9] = {0.125,
                              -0.807291666666666666666666666667e-1,
                              0.246028645833333333333333333333,
                              -1.82443876720610119047619047619,
                              25.3364147973439050099206349206,
                              -567.644412135183381139802038240,
                              18690.4765282320653831636345064,
                              -8.49353580299148769921876983660e5,
                              5.09225462402226769498681286758e7};
  double acc = m[7] + m[8] * r2;
  for (int i = 6; i >= 0; --i) acc = m[i] + r2 * acc;
  return ak + r * acc;
}

// gauss_legendre.rs:5713-5730 — J1(j0k)^2: table for k < 21, asymptotic series otherwise.
double bessel_j1_squared(size_t k) {
  if (k < 21) return dzb_gl_tables::J1[k - 1];
  const double x = 1.0 / ((double)k - 0.25);
  const double x2 = x * x;
  static const double s[8] = {-0.303380429711290253026202643516e-3, 0.198924364245969295201137972743e-3,
                              -0.228969902772111653038747229723e-3, 0.433710719130746277915572905025e-3,
                              -0.123632349727175414724737657367e-2, 0.496101423268883102872271417616e-2,
                              -0.266837393702323757700998557826e-1, 0.185395398206345628711318848386};
  double acc = s[6] + s[7] * x2;
  for (int i = 5; i >= 0; --i) acc = s[i] + x2 * acc;
  return x * (0.202642367284675542887758926420 + x2 * x2 * acc);
}

// gauss_legendre.rs:5743-5853 — Bogaert's asymptotic (theta, weight) for n >= 101.
void gauss_legendre_quad_pair_calculated(size_t n, size_t k, double* theta_out, double* w_out) {
  const double w = 1.0 / ((double)n + 0.5);
  const double nu = bessel_j_zero(k);
  double theta = w * nu;
  const double x = theta * theta;
  const double b = bessel_j1_squared(k);

  // Chebyshev interpolants for the nodes (:5754-5786) ...
  static const double SF1[7] = {-1.29052996274280508473467968379e-12, 2.40724685864330121825976175184e-10,
                                -3.13148654635992041468855740012e-8,  0.275573168962061235623801563453e-5,
                                -0.148809523713909147898955880165e-3, 0.416666666665193394525296923981e-2,
                                -0.416666666666662959639712457549e-1};
  static const double SF2[7] = {2.20639421781871003734786884322e-9,  -7.53036771373769326811030753538e-8,
                                0.161969259453836261731700382098e-5, -0.253300326008232025914059965302e-4,
                                0.282116886057560434805998583817e-3, -0.209022248387852902722635654229e-2,
                                0.815972221772932265640401128517e-2};
  static const double SF3[7] = {-2.97058225375526229899781956673e-8, 5.55845330223796209655886325712e-7,
                                -0.567797841356833081642185432056e-5, 0.418498100329504574443885193835e-4,
                                -0.251395293283965914823026348764e-3, 0.128654198542845137196151147483e-2,
                                -0.416012165620204364833694266818e-2};
  // ... and for the weights (:5789-5832)
  static const double WSF1[10] = {-2.20902861044616638398573427475e-14, 2.30365726860377376873232578871e-12,
                                  -1.75257700735423807659851042318e-10, 1.03756066927916795821098009353e-8,
                                  -4.63968647553221331251529631098e-7,  0.149644593625028648361395938176e-4,
                                  -0.326278659594412170300449074873e-3, 0.436507936507598105249726413120e-2,
                                  -0.305555555555553028279487898503e-1, 0.833333333333333302184063103900e-1};
  static const double WSF2[9] = {3.63117412152654783455929483029e-12, 7.67643545069893130779501844323e-11,
                                 -7.12912857233642220650643150625e-9, 2.11483880685947151466370130277e-7,
                                 -0.381817918680045468483009307090e-5, 0.465969530694968391417927388162e-4,
                                 -0.407297185611335764191683161117e-3, 0.268959435694729660779984493795e-2,
                                 -0.111111111111214923138249347172e-1};
  static const double WSF3[9] = {2.01826791256703301806643264922e-9,  -4.38647122520206649251063212545e-8,
                                 5.08898347288671653137451093208e-7,  -0.397933316519135275712977531366e-5,
                                 0.200559326396458326778521795392e-4, -0.422888059282921161626339411388e-4,
                                 -0.105646050254076140548678457002e-3, -0.947969308958577323145923317955e-4,
                                 0.656966489926484797412985260842e-2};
  const double sf1t = horner(SF1, 7, x), sf2t = horner(SF2, 7, x), sf3t = horner(SF3, 7, x);
  const double wsf1t = horner(WSF1, 10, x), wsf2t = horner(WSF2, 9, x), wsf3t = horner(WSF3, 9, x);

  // refinement with the paper's expansions (:5835-5852)
  const double nuosin = nu / std::sin(theta);
  const double bnuosin = b * nuosin;
  const double winvsinc = w * w * nuosin;
  const double wis2 = winvsinc * winvsinc;
  theta = w * (nu + theta * winvsinc * (sf1t + wis2 * (sf2t + wis2 * sf3t)));
  const double deno = bnuosin + bnuosin * wis2 * (wsf1t + wis2 * (wsf2t + wis2 * wsf3t));
  *theta_out = theta;
  *w_out = (2.0 * w) / deno;
}

// gauss_legendre.rs:5864-5896 — tabulated pair, 0-based k.  Row m-1 of the *_THETA_ZEROS / *_WEIGHTS
// tables holds m entries in decreasing theta; our generated header stores the rows back to back.
void gauss_legendre_quad_pair_tabulated(size_t l, size_t k, double* theta, double* w) {
  using namespace dzb_gl_tables;
  if (l & 1) {
    const size_t l2 = (l - 1) / 2;
    if (k == l2) {
      *theta = PI / 2.0;
      *w = 2.0 / (CL[l] * CL[l]);
    } else if (k < l2) {
      const size_t idx = ODD_OFF[l2 - 1] + (l2 - k - 1);
      *theta = ODD_THETA[idx];
      *w = ODD_W[idx];
    } else {
      const size_t idx = ODD_OFF[l2 - 1] + (k - l2 - 1);
      *theta = PI - ODD_THETA[idx];
      *w = ODD_W[idx];
    }
  } else {
    const size_t l2 = l / 2;
    if (k < l2) {
      const size_t idx = EVEN_OFF[l2 - 1] + (l2 - k - 1);
      *theta = EVEN_THETA[idx];
      *w = EVEN_W[idx];
    } else {
      const size_t idx = EVEN_OFF[l2 - 1] + (k - l2);
      *theta = PI - EVEN_THETA[idx];
      *w = EVEN_W[idx];
    }
  }
}

// gauss_legendre.rs:5908-5927 — 1-based k, valid only for k < n (the k == n node is unreachable).
// k == 0 would underflow `k - 1` in the reference (panic); reported as INTEGRATION here.
int quad_pair(size_t n, size_t k, double* theta, double* w) {
  if (!(k < n) || k == 0) return INTEGRATION;
  if (n < 101) {
    gauss_legendre_quad_pair_tabulated(n, k - 1, theta, w);
  } else if (2 * k - 1 > n) {
    gauss_legendre_quad_pair_calculated(n, n - k + 1, theta, w);
    *theta = PI - *theta;
  } else {
    gauss_legendre_quad_pair_calculated(n, k, theta, w);
  }
  return OK;
}

// The (x_j, w_j), j = 1..G-1, that every integration loop of the reference consumes
// (time_independent.rs:132-135 et al.): x_j = cos(theta_j) through libm.
static int quad_table(size_t G, std::vector<double>* xs, std::vector<double>* ws) {
  xs->clear();
  ws->clear();
  for (size_t j = 1; j < G; ++j) {
    double th, w;
    int rc = quad_pair(G, j, &th, &w);
    if (rc) return rc;
    xs->push_back(std::cos(th));
    ws->push_back(w);
  }
  return OK;
}

// ============================================================== basis
FirstDegreePolynomial FirstDegreePolynomial::transformation_to_0_1(double beg, double end) {  // polynomials_1d.rs:64-71
  return {1.0 / (end - beg), -beg / (end - beg)};
}
FirstDegreePolynomial FirstDegreePolynomial::transformation_from_m1_p1(double beg, double end) {  // :74-81
  return {(end - beg) / 2.0, (end + beg) / 2.0};
}

int PiecewiseFirstDegreePolynomial::from_polynomials(const FirstDegreePolynomial* p, int np, const double* bp, int nbp,
                                                     PiecewiseFirstDegreePolynomial* out) {
  if (np != nbp + 1 || np > 4) return PIECEWISE_DIMS;  // piecewise_polynomials_1degree.rs:112-114
  out->n_poly = np;
  for (int i = 0; i < np; ++i) out->polynomials[i] = p[i];
  for (int i = 0; i < nbp; ++i) out->interval_breakpoints[i] = bp[i];
  return OK;
}
bool PiecewiseFirstDegreePolynomial::operator==(const PiecewiseFirstDegreePolynomial& o) const {
  if (n_poly != o.n_poly) return false;
  for (int i = 0; i < n_poly; ++i)
    if (!(polynomials[i] == o.polynomials[i])) return false;
  for (int i = 0; i + 1 < n_poly; ++i)
    if (interval_breakpoints[i] != o.interval_breakpoints[i]) return false;
  return true;
}

// linear_basis.rs:31-94
PiecewiseFirstDegreePolynomial basis_function(const double* mesh, size_t n, size_t i) {
  typedef FirstDegreePolynomial P;
  PiecewiseFirstDegreePolynomial f;
  if (i == 0) {  // :33-44  [0, phi_2 o T(m0,m1), 0], breakpoints [m0, m1]
    P t = P::transformation_to_0_1(mesh[0], mesh[1]);
    P pieces[3] = {P::zero(), P::phi_2().compose(t), P::zero()};
    double bp[2] = {mesh[0], mesh[1]};
    PiecewiseFirstDegreePolynomial::from_polynomials(pieces, 3, bp, 2, &f);
  } else if (i == n - 1) {  // :74-89  [0, phi_1 o T(m_{n-2}, m_{n-1}), 0]
    P t = P::transformation_to_0_1(mesh[n - 2], mesh[n - 1]);
    P pieces[3] = {P::zero(), P::phi_1().compose(t), P::zero()};
    double bp[2] = {mesh[n - 2], mesh[n - 1]};
    PiecewiseFirstDegreePolynomial::from_polynomials(pieces, 3, bp, 2, &f);
  } else {  // :49-71  [0, phi_1 o T(prev,cur), phi_2 o T(cur,next), 0], breakpoints [prev, cur, next]
    const double prev = mesh[i - 1], cur = mesh[i], next = mesh[i + 1];
    P tl = P::transformation_to_0_1(prev, cur);
    P left = P::phi_1().compose(tl);
    P tr = P::transformation_to_0_1(cur, next);
    P right = P::phi_2().compose(tr);
    P pieces[4] = {P::zero(), left, right, P::zero()};
    double bp[3] = {prev, cur, next};
    PiecewiseFirstDegreePolynomial::from_polynomials(pieces, 4, bp, 3, &f);
  }
  return f;
}
std::vector<PiecewiseFirstDegreePolynomial> linear_basis_new(const double* mesh, size_t n) {
  std::vector<PiecewiseFirstDegreePolynomial> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) v.push_back(basis_function(mesh, n, i));
  return v;
}

// ============================================================== linear algebra
// matrix_solver/mod.rs:17-48.  (The reference panics for n < 2; reported as WRONG_DIMS.)
template <class Mat>
int solve_by_thomas(const Mat& m, const double* b, size_t nb, double* solution) {
  if (m.n != nb || nb < 2) return WRONG_DIMS;  // :19-21
  const size_t n = nb;
  std::vector<double> c(n - 1, 0.0), d(n, 0.0);
  c[0] = m.get(0, 1) / m.get(0, 0);  // :27
  d[0] = b[0] / m.get(0, 0);         // :28
  for (size_t i = 1; i < n - 1; ++i) {  // :30-35
    c[i] = m.get(i, i + 1) / (m.get(i, i) - m.get(i, i - 1) * c[i - 1]);
    d[i] = (b[i] - m.get(i, i - 1) * d[i - 1]) / (m.get(i, i) - m.get(i, i - 1) * c[i - 1]);
  }
  d[n - 1] = (b[n - 1] - m.get(n - 1, n - 2) * d[n - 2]) /
             (m.get(n - 1, n - 1) - m.get(n - 1, n - 2) * c[n - 2]);  // :37-39
  solution[n - 1] = d[n - 1];                                          // :41
  for (size_t i = n - 1; i-- > 0;) solution[i] = d[i] - c[i] * solution[i + 1];  // :43-45
  return OK;
}

// fem/utils.rs:41-62
template <class Mat>
int tridiagonal_matrix_vector_multiplication(const Mat& a, const double* b, size_t nb, double c, double* out) {
  if (a.n != nb || nb < 2) return WRONG_DIMS;  // :43-45
  const size_t len = nb;
  for (size_t i = 1; i + 1 < len; ++i)  // :52-56
    out[i] = c * (a.get(i, i - 1) * b[i - 1] + a.get(i, i) * b[i] + a.get(i, i + 1) * b[i + 1]);
  out[0] = c * (a.get(0, 0) * b[0] + a.get(0, 1) * b[1]);                                      // :58
  out[len - 1] = c * (a.get(len - 1, len - 2) * b[len - 2] + a.get(len - 1, len - 1) * b[len - 1]);  // :59
  return OK;
}
// fem/utils.rs:16-30
int add(const double* b, size_t nb, const double* v, size_t nv, double* out) {
  if (nb != nv) return WRONG_DIMS;
  for (size_t i = 0; i < nb; ++i) out[i] = b[i] + v[i];
  return OK;
}

// ============================================================== assembly
struct QuadSource {  // hoisted table or per-(i,j) recomputation, same values either way
  size_t G;
  bool hoist;
  std::vector<double> xs, ws;
  int init(size_t G_, bool hoist_) {
    G = G_;
    hoist = hoist_;
    // the reference only discovers a bad (G, j) inside the loop; with a mesh of >= 3 nodes the first
    // iteration j = 1 fails iff G < 2 — and G <= 1 means the loop `1..G` is empty, so no error at all.
    if (hoist && G >= 2) return quad_table(G, &xs, &ws);
    return OK;
  }
  inline int get(size_t j, double* x, double* w) const {
    if (hoist) {
      *x = xs[j - 1];
      *w = ws[j - 1];
      return OK;
    }
    double th;
    int rc = quad_pair(G, j, &th, w);  // time_independent.rs:134
    if (rc) return rc;
    *x = std::cos(th);  // :135
    return OK;
  }
};

// diffusion_solver/time_independent.rs:91-182
template <class Mat>
int ti_gauss_legendre_integration(const double bc[2], double mu, double b, const double* mesh, size_t n,
                                  size_t gauss_step, bool hoist_table, int threads, Mat* K, double* b_vector) {
  typedef FirstDegreePolynomial P;
  if (n < 3 || K->n != n) return WRONG_DIMS;  // the reference indexes mesh[1], basis[i+1]: n < 2 panics
  QuadSource q;
  int rc0 = q.init(gauss_step, hoist_table);
  if (rc0) return rc0;
  for (size_t i = 0; i < n; ++i) b_vector[i] = 0.0;  // :99
  int err = OK;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) if (threads > 1)
  for (long long ii = 1; ii < (long long)n - 1; ++ii) {  // :102
    const size_t i = (size_t)ii;
    const PiecewiseFirstDegreePolynomial phi = basis_function(mesh, n, i);
    const PiecewiseFirstDegreePolynomial derivative_phi = phi.differentiate();                   // :104
    const P t_prev = P::transformation_from_m1_p1(mesh[i - 1], mesh[i]);                         // :106-109
    const P t_next = P::transformation_from_m1_p1(mesh[i], mesh[i + 1]);                         // :110-113
    const P t_square = P::transformation_from_m1_p1(mesh[i - 1], mesh[i + 1]);                   // :114-118
    const P dt_prev = t_prev.differentiate(), dt_next = t_next.differentiate(), dt_square = t_square.differentiate();
    const PiecewiseFirstDegreePolynomial derivative_prev = basis_function(mesh, n, i - 1).differentiate();  // :124
    const PiecewiseFirstDegreePolynomial derivative_next = basis_function(mesh, n, i + 1).differentiate();  // :125
    double integral_prev = 0.0, integral_next = 0.0, integral_square = 0.0;
    for (size_t j = 1; j < gauss_step; ++j) {  // :132
      double x, w;
      int rc = q.get(j, &x, &w);
      if (rc) {
        err = rc;
        break;
      }
      const double tp_prev = t_prev.evaluate(x), tp_next = t_next.evaluate(x), tp_square = t_square.evaluate(x);
      integral_prev += (mu * derivative_phi.evaluate(tp_prev) * derivative_prev.evaluate(tp_prev) +
                        b * derivative_prev.evaluate(tp_prev) * phi.evaluate(tp_prev)) *
                       dt_prev.evaluate(x) * w;  // :142-149
      integral_next += (mu * derivative_phi.evaluate(tp_next) * derivative_next.evaluate(tp_next) +
                        b * derivative_next.evaluate(tp_next) * phi.evaluate(tp_next)) *
                       dt_next.evaluate(x) * w;  // :150-157
      integral_square += (mu * derivative_phi.evaluate(tp_square) * derivative_phi.evaluate(tp_square) +
                          b * derivative_phi.evaluate(tp_square) * phi.evaluate(tp_square)) *
                         dt_square.evaluate(x) * w;  // :158-165
    }
    K->set(i, i, integral_square);    // :168
    K->set(i, i - 1, integral_prev);  // :169
    K->set(i, i + 1, integral_next);  // :170
  }
  if (err) return err;
  K->set(0, 0, 1.0);          // :176
  K->set(n - 1, n - 1, 1.0);  // :177
  b_vector[0] = bc[0];        // :178
  b_vector[n - 1] = bc[1];    // :179
  return OK;
}

// diffusion_solver/time_dependent.rs:110-247.  `.powf(2_f64)` (:185,:199) is restated as v*v: LLVM folds
// pow(x, 2.0) into x*x in optimised builds, and glibc's pow agrees with x*x except on rare near-ties.
template <class Mat>
int td_gauss_legendre_integration(double mu, double b, const double* mesh, size_t n, size_t gauss_step,
                                  bool hoist_table, int threads, Mat* M, Mat* S) {
  typedef FirstDegreePolynomial P;
  if (n < 3 || M->n != n || S->n != n) return WRONG_DIMS;
  QuadSource q;
  int rc0 = q.init(gauss_step, hoist_table);
  if (rc0) return rc0;
  int err = OK;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) if (threads > 1)
  for (long long ii = 1; ii < (long long)n - 1; ++ii) {  // :121
    const size_t i = (size_t)ii;
    const PiecewiseFirstDegreePolynomial phi = basis_function(mesh, n, i);
    const PiecewiseFirstDegreePolynomial phi_prev = basis_function(mesh, n, i - 1);
    const PiecewiseFirstDegreePolynomial phi_next = basis_function(mesh, n, i + 1);
    const PiecewiseFirstDegreePolynomial derivative_phi = phi.differentiate();            // :125
    const PiecewiseFirstDegreePolynomial derivative_phi_next = phi_next.differentiate();  // :127
    const PiecewiseFirstDegreePolynomial derivative_phi_prev = phi_prev.differentiate();  // :129
    const P t_prev = P::transformation_from_m1_p1(mesh[i - 1], mesh[i]);                  // :132-135
    const P t_next = P::transformation_from_m1_p1(mesh[i], mesh[i + 1]);                  // :136-139
    const P t_square = P::transformation_from_m1_p1(mesh[i - 1], mesh[i + 1]);            // :140-144
    const P dt_prev = t_prev.differentiate(), dt_next = t_next.differentiate(), dt_square = t_square.differentiate();
    double prev_prime = 0.0, next_prime = 0.0, square_prime = 0.0;  // <phi_j', phi_i'>   :153-155
    double prev_half = 0.0, next_half = 0.0, square_half = 0.0;     // <phi_j, phi_i'>    :157-159
    double prev_mass = 0.0, next_mass = 0.0, square_mass = 0.0;     // <phi_j, phi_i>     :161-163
    for (size_t j = 1; j < gauss_step; ++j) {                       // :166
      double x, w;
      int rc = q.get(j, &x, &w);
      if (rc) {
        err = rc;
        break;
      }
      const double tp_prev = t_prev.evaluate(x), tp_next = t_next.evaluate(x), tp_square = t_square.evaluate(x);
      prev_mass += phi.evaluate(tp_prev) * phi_prev.evaluate(tp_prev) * dt_prev.evaluate(x) * w;  // :180-182
      {
        const double v = phi.evaluate(tp_square);
        square_mass += (v * v) * dt_square.evaluate(x) * w;  // :184-186
      }
      next_mass += phi.evaluate(tp_next) * phi_next.evaluate(tp_next) * dt_next.evaluate(x) * w;  // :188-190
      prev_prime += derivative_phi.evaluate(tp_prev) * derivative_phi_prev.evaluate(tp_prev) * dt_prev.evaluate(x) * w;  // :194-196
      {
        const double v = derivative_phi.evaluate(tp_square);
        square_prime += (v * v) * dt_square.evaluate(x) * w;  // :198-200
      }
      next_prime += derivative_phi.evaluate(tp_next) * derivative_phi_next.evaluate(tp_next) * dt_next.evaluate(x) * w;  // :202-204
      prev_half += phi.evaluate(tp_prev) * derivative_phi_prev.evaluate(tp_prev) * dt_prev.evaluate(x) * w;        // :208-210
      square_half += phi.evaluate(tp_square) * derivative_phi.evaluate(tp_square) * dt_square.evaluate(x) * w;     // :212-214
      next_half += phi.evaluate(tp_next) * derivative_phi_next.evaluate(tp_next) * dt_next.evaluate(x) * w;        // :216-218
    }
    M->set(i, i - 1, prev_mass);    // :221
    M->set(i, i, square_mass);      // :222
    M->set(i, i + 1, next_mass);    // :223
    S->set(i, i - 1, -mu * prev_prime - b * prev_half);      // :226-227
    S->set(i, i, -mu * square_prime - b * square_half);      // :229-230
    S->set(i, i + 1, -mu * next_prime - b * next_half);      // :232-233
  }
  if (err) return err;
  M->set(0, 0, 1.0);          // :236
  M->set(n - 1, n - 1, 1.0);  // :237
  S->set(0, 0, 1.0);          // :238
  S->set(n - 1, n - 1, 1.0);  // :239
  return OK;
}

// stokes_solver/dim1.rs:114-240.  The closure is called in the reference's order: interior rows
// i = 1..n-2 (j ascending inside), then row 0.
template <class Mat>
int stokes_gauss_legendre_integration(double rho, double hydrostatic_pressure, const double* mesh, size_t n,
                                      size_t gauss_step, force_fn function, void* ctx, bool hoist_table, Mat* K,
                                      double* b_vector) {
  typedef FirstDegreePolynomial P;
  if (n < 3 || K->n != n) return WRONG_DIMS;
  QuadSource q;
  int rc0 = q.init(gauss_step, hoist_table);
  if (rc0) return rc0;
  for (size_t i = 0; i < n; ++i) b_vector[i] = 0.0;  // :122
  for (size_t i = 1; i + 1 < n; ++i) {               // :125
    const PiecewiseFirstDegreePolynomial phi = basis_function(mesh, n, i);
    const PiecewiseFirstDegreePolynomial derivative_phi = phi.differentiate();       // :127
    const P t_prev = P::transformation_from_m1_p1(mesh[i - 1], mesh[i]);             // :129-132
    const P t_next = P::transformation_from_m1_p1(mesh[i], mesh[i + 1]);             // :133-136
    const P t_square = P::transformation_from_m1_p1(mesh[i - 1], mesh[i + 1]);       // :137-141
    const P dt_prev = t_prev.differentiate(), dt_next = t_next.differentiate(), dt_square = t_square.differentiate();
    const PiecewiseFirstDegreePolynomial derivative_prev = basis_function(mesh, n, i - 1).differentiate();  // :147
    const PiecewiseFirstDegreePolynomial derivative_next = basis_function(mesh, n, i + 1).differentiate();  // :148
    double integral_prev = 0.0, integral_next = 0.0, integral_square = 0.0, b_integral = 0.0;
    for (size_t j = 1; j < gauss_step; ++j) {  // :156
      double x, w;
      int rc = q.get(j, &x, &w);
      if (rc) return rc;
      const double tp_prev = t_prev.evaluate(x), tp_next = t_next.evaluate(x), tp_square = t_square.evaluate(x);
      integral_prev += phi.evaluate(tp_prev) * derivative_prev.evaluate(tp_prev) * dt_prev.evaluate(x) * w;        // :166-170
      integral_next += phi.evaluate(tp_next) * derivative_next.evaluate(tp_next) * dt_next.evaluate(x) * w;        // :171-175
      integral_square += phi.evaluate(tp_square) * derivative_phi.evaluate(tp_square) * dt_square.evaluate(x) * w;  // :176-180
      b_integral += rho * function(tp_square, ctx) * phi.evaluate(tp_square) * dt_square.evaluate(x) * w;         // :181-185
    }
    K->set(i, i, integral_square);    // :188
    K->set(i, i - 1, integral_prev);  // :189
    K->set(i, i + 1, integral_next);  // :190
    b_vector[i] = b_integral;         // :191
  }
  const PiecewiseFirstDegreePolynomial phi0 = basis_function(mesh, n, 0);
  const PiecewiseFirstDegreePolynomial derivative_phi_0 = phi0.differentiate();                          // :195
  const PiecewiseFirstDegreePolynomial derivative_phi_1 = basis_function(mesh, n, 1).differentiate();    // :196
  const P t_square_0 = P::transformation_from_m1_p1(mesh[0], mesh[1]);                                    // :198-202
  const P dt_square_0 = t_square_0.differentiate();
  double integral_0 = 0.0, integral_0_next = 0.0, b_first = 0.0;
  for (size_t j = 1; j < gauss_step; ++j) {  // :210
    double x, w;
    int rc = q.get(j, &x, &w);
    if (rc) return rc;
    const double t0 = t_square_0.evaluate(x);
    integral_0 += phi0.evaluate(t0) * derivative_phi_0.evaluate(t0) * dt_square_0.evaluate(x) * w;       // :218-220
    integral_0_next += phi0.evaluate(t0) * derivative_phi_1.evaluate(t0) * dt_square_0.evaluate(x) * w;  // :222-224
    b_first += rho * function(t0, ctx) * phi0.evaluate(t0) * dt_square_0.evaluate(x) * w;                // :226-228
  }
  K->set(0, 0, integral_0);             // :232
  K->set(0, 1, integral_0_next);        // :233
  K->set(n - 1, n - 1, 1.0);            // :234
  b_vector[0] = b_first;                // :235
  b_vector[n - 1] = hydrostatic_pressure;  // :236
  return OK;
}

// diffusion_solver/time_dependent.rs:257-280
template <class Mat>
int td_solve(const Mat& mass, const Mat& stiff, double* state, size_t n, double time_step, const double bc[2]) {
  std::vector<double> b1(n), b2(n), b(n), res(n);
  int rc = tridiagonal_matrix_vector_multiplication(stiff, state, n, time_step, b1.data());  // :260-261
  if (rc) return rc;
  rc = tridiagonal_matrix_vector_multiplication(mass, state, n, 1.0, b2.data());  // :263-264
  if (rc) return rc;
  rc = add(b1.data(), n, b2.data(), n, b.data());  // :266-268
  if (rc) return rc;
  rc = solve_by_thomas(mass, b.data(), n, res.data());  // :270
  if (rc) return rc;
  res[0] = bc[0];      // :273
  res[n - 1] = bc[1];  // :274
  std::memcpy(state, res.data(), n * sizeof(double));  // :276
  return OK;
}

// ============================================================== mesh ingestion
// Rust `str::parse::<f64>` grammar (core::num::dec2flt): [+-] ( inf | infinity | nan | digits[.digits][e[+-]digits] ),
// case-insensitive words, at least one digit in the mantissa, nothing else (no blanks, no hex).  Values are
// correctly rounded, as glibc strtod is.
static bool rust_parse_f64(const std::string& s, double* out) {
  size_t p = 0;
  const size_t n = s.size();
  if (n == 0) return false;
  if (s[p] == '+' || s[p] == '-') ++p;
  if (p >= n) return false;
  auto ieq = [&](const char* wd) {
    size_t L = std::strlen(wd);
    if (n - p != L) return false;
    for (size_t i = 0; i < L; ++i)
      if (std::tolower((unsigned char)s[p + i]) != wd[i]) return false;
    return true;
  };
  if (ieq("inf") || ieq("infinity") || ieq("nan")) {
    *out = std::strtod(s.c_str(), nullptr);
    return true;
  }
  size_t digits = 0;
  while (p < n && std::isdigit((unsigned char)s[p])) ++p, ++digits;
  if (p < n && s[p] == '.') {
    ++p;
    while (p < n && std::isdigit((unsigned char)s[p])) ++p, ++digits;
  }
  if (digits == 0) return false;
  if (p < n && (s[p] == 'e' || s[p] == 'E')) {
    ++p;
    if (p < n && (s[p] == '+' || s[p] == '-')) ++p;
    size_t ed = 0;
    while (p < n && std::isdigit((unsigned char)s[p])) ++p, ++ed;
    if (ed == 0) return false;
  }
  if (p != n) return false;
  *out = std::strtod(s.c_str(), nullptr);
  return true;
}
static void split_single_space(const std::string& line, std::vector<std::string>* parts) {  // str::split(" ")
  parts->clear();
  size_t start = 0;
  for (;;) {
    size_t pos = line.find(' ', start);
    if (pos == std::string::npos) {
      parts->push_back(line.substr(start));
      break;
    }
    parts->push_back(line.substr(start, pos - start));
    start = pos + 1;
  }
}
static bool read_lines(const char* path, std::vector<std::string>* lines) {  // BufRead::lines(): strips \n and \r\n
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::string l;
  while (std::getline(f, l)) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    lines->push_back(l);
  }
  return true;
}

struct Mesh1D {
  std::vector<double> vertices;   // 12 n
  std::vector<uint32_t> indices;  // 6 n - 6
  double max_length = 0, prom_width = 0;
  float model_translation[3] = {0, 0, 0};
};

// mesh/mesh_builder.rs:140-190 (constant-axis detection) + :204-328 (build_mesh_1d) + :58-84 (vertex check)
static int build_mesh_1d(const char* path, bool has_mult, double height_multiplier, Mesh1D* m) {
  std::vector<std::string> lines;
  if (!read_lines(path, &lines)) return IO;
  std::vector<std::string> parts;
  // ---- check_for_constant_coordinates :140-190
  std::set<std::string> keys[3];
  for (const std::string& line : lines) {
    if (line.rfind("v ", 0) != 0) continue;  // starts_with("v ") :159
    split_single_space(line, &parts);
    if (parts.size() - 1 != 3) {
      // parse errors surface first (the map runs before try_into :175-178)
      for (size_t t = 1; t < parts.size(); ++t) {
        double tmp;
        if (!rust_parse_f64(parts[t], &tmp)) return MESH_PARSE;
      }
      return MESH_PARSE;  // Error::Infallible from try_into
    }
    for (int a = 0; a < 3; ++a) {
      const std::string& c = parts[a + 1];
      double tmp;
      if (!rust_parse_f64(c, &tmp)) return MESH_PARSE;  // c_str.parse::<f32>()? :168-170 (same grammar as f64)
      if (c.rfind("0.0", 0) == 0 || c.rfind("-0.0", 0) == 0)
        keys[a].insert("0.0");  // :167-168
      else
        keys[a].insert(c);
    }
  }
  // ---- constant axis pair :217-227
  int cc[2];
  if (keys[0].size() == 1 && keys[1].size() == 1) {
    cc[0] = 1, cc[1] = 0;
  } else if (keys[1].size() == 1 && keys[2].size() == 1) {
    cc[0] = 2, cc[1] = 1;
  } else if (keys[2].size() == 1 && keys[0].size() == 1) {
    cc[0] = 2, cc[1] = 0;
  } else {
    return MESH_PARSE;
  }
  // ---- ordered vertices :230-266
  std::vector<double>& vertices = m->vertices;
  vertices.clear();
  for (const std::string& line : lines) {
    if (line.rfind("v ", 0) != 0) continue;
    split_single_space(line, &parts);  // obj_vertex_checker :58-84
    std::vector<double> coordinate;
    for (size_t t = 1; t < parts.size(); ++t) {
      double v;
      if (!rust_parse_f64(parts[t], &v)) return MESH_PARSE;
      coordinate.push_back(v);
    }
    if (coordinate.size() != 3) return MESH_PARSE;
    for (int k = 0; k < 2; ++k) {  // :244-247
      coordinate.erase(coordinate.begin() + cc[k]);
      coordinate.push_back(0.0);
    }
    const double new_value = coordinate[0];
    vertices.insert(vertices.end(), coordinate.begin(), coordinate.end());
    vertices.push_back(0.0), vertices.push_back(0.0), vertices.push_back(1.0);  // :253
    long long j = (long long)vertices.size() - 6 - 5 - 1;                       // :255
    while (j >= 0 && vertices[(size_t)j] > new_value) {                          // :256-259
      vertices[(size_t)j + 6] = vertices[(size_t)j];
      j -= 6;
    }
    vertices[(size_t)(j + 6)] = new_value;  // :260
  }
  if (vertices.empty()) return MESH_PARSE;  // the reference panics on vertices[0]
  const uint32_t vertices_len = (uint32_t)vertices.size();
  m->max_length = -vertices[0] + vertices[vertices_len - 6];  // :270
  const double multiplier = has_mult ? height_multiplier : 1.0;
  m->prom_width = (m->max_length * 6.0 / ((double)vertices_len - 6.)) * multiplier;  // :277
  {
    std::vector<double> copy(vertices);  // :279-286
    for (size_t idx = 0; idx < copy.size(); ++idx)
      if (idx % 6 == 1) copy[idx] = m->prom_width;
    vertices.insert(vertices.end(), copy.begin(), copy.end());
  }
  std::vector<uint32_t>& ind = m->indices;
  ind.clear();
  ind.push_back(0), ind.push_back(1), ind.push_back(vertices_len / 6);  // :290
  ind.push_back((vertices_len - 6) / 6), ind.push_back((vertices_len * 2 - 6) / 6),
      ind.push_back((vertices_len * 2 - 12) / 6);  // :292-296
  for (uint32_t i = 1; i + 1 < vertices_len / 6; ++i) {  // :298-307
    ind.push_back(i), ind.push_back(i + vertices_len / 6), ind.push_back(i + vertices_len / 6 - 1);
    ind.push_back(i), ind.push_back(i + 1), ind.push_back(i + vertices_len / 6);
  }
  float middle_point[3] = {0.f, 0.f, 0.f};
  middle_point[0] = (float)vertices[0] + (float)m->max_length / 2.0f;  // :310
  middle_point[1] = (float)m->prom_width / 2.0f;                       // :311
  m->model_translation[0] = -middle_point[0];                          // :314-318
  m->model_translation[1] = -middle_point[1];
  m->model_translation[2] = middle_point[2];
  return OK;
}

}  // namespace dzo

// ================================================================== C interface for ctypes (tests, bench)
using namespace dzo;

namespace {
template <class Mat>
void extract_bands(const Mat& m, double* lo, double* di, double* up) {
  const size_t n = m.n;
  for (size_t i = 0; i < n; ++i) {
    lo[i] = i ? m.get(i, i - 1) : 0.0;
    di[i] = m.get(i, i);
    up[i] = i + 1 < n ? m.get(i, i + 1) : 0.0;
  }
}
// In dense mode, everything outside the three diagonals must still be the zero fill.
bool dense_is_tridiagonal(const Dense& m) {
  for (size_t i = 0; i < m.n; ++i)
    for (size_t j = 0; j < m.n; ++j)
      if ((j + 1 < i || j > i + 1) && m.get(i, j) != 0.0) return false;
  return true;
}
Banded from_bands(const double* lo, const double* di, const double* up, size_t n) {
  Banded m(n);
  for (size_t i = 0; i < n; ++i) m.lo[i] = lo[i], m.di[i] = di[i], m.up[i] = up[i];
  return m;
}
}  // namespace

extern "C" {

int dzo_hardware_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int dzo_quad_pair(size_t n, size_t k, double* theta, double* w) { return quad_pair(n, k, theta, w); }

// x[j-1] = cos(theta_j), w[j-1] for j = 1..G-1
int dzo_quad_table(size_t G, double* x, double* w) {
  std::vector<double> xs, ws;
  int rc = quad_table(G, &xs, &ws);
  if (rc) return rc;
  std::memcpy(x, xs.data(), xs.size() * sizeof(double));
  std::memcpy(w, ws.data(), ws.size() * sizeof(double));
  return OK;
}

// out: per basis function 4 coefficients, 4 independent terms, 3 breakpoints, n_poly (as double) = 12 doubles
int dzo_linear_basis(const double* mesh, size_t n, double* out) {
  if (n < 2) return WRONG_DIMS;
  std::vector<PiecewiseFirstDegreePolynomial> b = linear_basis_new(mesh, n);
  for (size_t i = 0; i < n; ++i) {
    double* o = out + 12 * i;
    for (int k = 0; k < 12; ++k) o[k] = 0.0;
    for (int k = 0; k < b[i].n_poly; ++k) o[k] = b[i].polynomials[k].coefficient, o[4 + k] = b[i].polynomials[k].independent_term;
    for (int k = 0; k + 1 < b[i].n_poly; ++k) o[8 + k] = b[i].interval_breakpoints[k];
    o[11] = (double)b[i].n_poly;
  }
  return OK;
}
double dzo_basis_evaluate(const double* mesh, size_t n, size_t i, double x, int derivative) {
  PiecewiseFirstDegreePolynomial f = basis_function(mesh, n, i);
  return derivative ? f.differentiate().evaluate(x) : f.evaluate(x);
}

// mode: 0 = banded storage + hoisted table (fast oracle), 1 = banded + per-(i,j) quad_pair/cos,
//       2 = dense n x n Array2 stand-in + per-(i,j) recomputation (the reference's actual cost profile)
int dzo_ti_assemble(const double* mesh, size_t n, double mu, double b, double bc0, double bc1, size_t G, int mode,
                    int threads, double* lo, double* di, double* up, double* rhs) {
  const double bc[2] = {bc0, bc1};
  if (mode == 2) {
    Dense K(n);
    int rc = ti_gauss_legendre_integration(bc, mu, b, mesh, n, G, false, threads, &K, rhs);
    if (rc) return rc;
    if (!dense_is_tridiagonal(K)) return WRONG_DIMS;
    extract_bands(K, lo, di, up);
    return OK;
  }
  Banded K(n);
  int rc = ti_gauss_legendre_integration(bc, mu, b, mesh, n, G, mode == 0, threads, &K, rhs);
  if (rc) return rc;
  extract_bands(K, lo, di, up);
  return OK;
}

int dzo_td_assemble(const double* mesh, size_t n, double mu, double b, size_t G, int mode, int threads, double* mlo,
                    double* mdi, double* mup, double* slo, double* sdi, double* sup) {
  if (mode == 2) {
    Dense M(n), S(n);
    int rc = td_gauss_legendre_integration(mu, b, mesh, n, G, false, threads, &M, &S);
    if (rc) return rc;
    if (!dense_is_tridiagonal(M) || !dense_is_tridiagonal(S)) return WRONG_DIMS;
    extract_bands(M, mlo, mdi, mup);
    extract_bands(S, slo, sdi, sup);
    return OK;
  }
  Banded M(n), S(n);
  int rc = td_gauss_legendre_integration(mu, b, mesh, n, G, mode == 0, threads, &M, &S);
  if (rc) return rc;
  extract_bands(M, mlo, mdi, mup);
  extract_bands(S, slo, sdi, sup);
  return OK;
}

int dzo_stokes_assemble(const double* mesh, size_t n, double rho, double hydrostatic_pressure, size_t G,
                        double (*f)(double, void*), void* ctx, int mode, double* lo, double* di, double* up,
                        double* rhs) {
  if (mode == 2) {
    Dense K(n);
    int rc = stokes_gauss_legendre_integration(rho, hydrostatic_pressure, mesh, n, G, f, ctx, false, &K, rhs);
    if (rc) return rc;
    if (!dense_is_tridiagonal(K)) return WRONG_DIMS;
    extract_bands(K, lo, di, up);
    return OK;
  }
  Banded K(n);
  int rc = stokes_gauss_legendre_integration(rho, hydrostatic_pressure, mesh, n, G, f, ctx, mode == 0, &K, rhs);
  if (rc) return rc;
  extract_bands(K, lo, di, up);
  return OK;
}

static double const_force(double, void* ctx) { return *(double*)ctx; }
int dzo_stokes_assemble_const_force(const double* mesh, size_t n, double rho, double hydrostatic_pressure, size_t G,
                                    double force, int mode, double* lo, double* di, double* up, double* rhs) {
  return dzo_stokes_assemble(mesh, n, rho, hydrostatic_pressure, G, const_force, &force, mode, lo, di, up, rhs);
}

int dzo_thomas(const double* lo, const double* di, const double* up, const double* rhs, size_t n, double* out) {
  Banded m = from_bands(lo, di, up, n);
  return solve_by_thomas(m, rhs, n, out);
}
// a: row-major n x n; `rows`/`cols` let the tests exercise the non-square WrongDims branch
int dzo_thomas_dense(const double* a, size_t rows, size_t cols, const double* b, size_t nb, double* out) {
  if (rows != cols) return WRONG_DIMS;  // matrix.is_square()  matrix_solver/mod.rs:19
  Dense m(rows);
  std::memcpy(m.a.data(), a, rows * cols * sizeof(double));
  return solve_by_thomas(m, b, nb, out);
}
int dzo_tridiag_matvec(const double* lo, const double* di, const double* up, const double* b, size_t n, size_t nb,
                       double c, double* out) {
  Banded m = from_bands(lo, di, up, n);
  return tridiagonal_matrix_vector_multiplication(m, b, nb, c, out);
}
int dzo_add(const double* b, size_t nb, const double* v, size_t nv, double* out) { return add(b, nb, v, nv, out); }

// `steps` successive solve(dt) calls on one bar; state updated in place (time_dependent.rs:257-280)
int dzo_td_run(const double* mlo, const double* mdi, const double* mup, const double* slo, const double* sdi,
               const double* sup, double* state, size_t n, double dt, double bc0, double bc1, size_t steps) {
  Banded M = from_bands(mlo, mdi, mup, n), S = from_bands(slo, sdi, sup, n);
  const double bc[2] = {bc0, bc1};
  for (size_t s = 0; s < steps; ++s) {
    int rc = td_solve(M, S, state, n, dt, bc);
    if (rc) return rc;
  }
  return OK;
}
// Same, on the dense Array2 stand-in (faithful cost profile; n <= 4096)
int dzo_td_run_dense(const double* mlo, const double* mdi, const double* mup, const double* slo, const double* sdi,
                     const double* sup, double* state, size_t n, double dt, double bc0, double bc1, size_t steps) {
  Dense M(n), S(n);
  for (size_t i = 0; i < n; ++i) {
    M.set(i, i, mdi[i]), S.set(i, i, sdi[i]);
    if (i) M.set(i, i - 1, mlo[i]), S.set(i, i - 1, slo[i]);
    if (i + 1 < n) M.set(i, i + 1, mup[i]), S.set(i, i + 1, sup[i]);
  }
  const double bc[2] = {bc0, bc1};
  for (size_t s = 0; s < steps; ++s) {
    int rc = td_solve(M, S, state, n, dt, bc);
    if (rc) return rc;
  }
  return OK;
}

// DiffussionSolverTimeDependent::new's state construction + IC length check (time_dependent.rs:62-78)
int dzo_td_initial_state(const double* ic, size_t n_ic, size_t n, double bc0, double bc1, double* state) {
  if (n < 2 || n_ic != n - 2) return WRONG_DIMS;  // :66-68
  for (size_t i = 0; i < n; ++i) state[i] = 0.0;
  state[0] = bc0;
  state[n - 1] = bc1;
  for (size_t i = 1; i + 1 < n; ++i) state[i] = ic[i - 1];
  return OK;
}

// ---- ensembles of independent bars (bar-major [B][n]); OpenMP over bars.  Used by the CPU baseline.
int dzo_ensemble_td_assemble(const double* meshes, size_t B, size_t n, double mu, double b, size_t G, int threads,
                             double* mlo, double* mdi, double* mup, double* slo, double* sdi, double* sup) {
  int err = OK;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1) if (threads > 1)
  for (long long bb = 0; bb < (long long)B; ++bb) {
    const size_t o = (size_t)bb * n;
    int rc = dzo_td_assemble(meshes + o, n, mu, b, G, 0, 1, mlo + o, mdi + o, mup + o, slo + o, sdi + o, sup + o);
    if (rc) err = rc;
  }
  return err;
}
int dzo_ensemble_td_run(const double* mlo, const double* mdi, const double* mup, const double* slo, const double* sdi,
                        const double* sup, double* state, size_t B, size_t n, double dt, double bc0, double bc1,
                        size_t steps, int threads) {
  int err = OK;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) if (threads > 1)
  for (long long bb = 0; bb < (long long)B; ++bb) {
    const size_t o = (size_t)bb * n;
    int rc = dzo_td_run(mlo + o, mdi + o, mup + o, slo + o, sdi + o, sup + o, state + o, n, dt, bc0, bc1, steps);
    if (rc) err = rc;
  }
  return err;
}

// ---- a resident ensemble for the timed CPU legs of bench.py: the matrices of every bar are built ONCE (what
// DiffussionSolverTimeDependent::new leaves in the struct, time_dependent.rs:49-58), the states live in the session and
// `advance` is `steps` x solve(dt) per bar inside one parallel region — no per-call band or state copies.
struct DzoEnsemble {
  size_t B = 0, n = 0;
  double bc[2] = {0, 0};
  std::vector<Banded> M, S;
  std::vector<double> state;
};
void* dzo_ensemble_td_open(const double* mlo, const double* mdi, const double* mup, const double* slo, const double* sdi,
                           const double* sup, const double* state, size_t B, size_t n, double bc0, double bc1) {
  DzoEnsemble* e = new DzoEnsemble;
  e->B = B, e->n = n, e->bc[0] = bc0, e->bc[1] = bc1;
  e->M.reserve(B), e->S.reserve(B);
  for (size_t b = 0; b < B; ++b) {
    const size_t o = b * n;
    e->M.push_back(from_bands(mlo + o, mdi + o, mup + o, n));
    e->S.push_back(from_bands(slo + o, sdi + o, sup + o, n));
  }
  e->state.assign(state, state + B * n);
  return e;
}
int dzo_ensemble_td_advance(void* h, double dt, size_t steps, int threads) {
  DzoEnsemble* e = (DzoEnsemble*)h;
  int err = OK;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) if (threads > 1)
  for (long long bb = 0; bb < (long long)e->B; ++bb) {
    double* st = e->state.data() + (size_t)bb * e->n;
    for (size_t s = 0; s < steps; ++s) {
      int rc = td_solve(e->M[(size_t)bb], e->S[(size_t)bb], st, e->n, dt, e->bc);
      if (rc) {
        err = rc;
        break;
      }
    }
  }
  return err;
}
int dzo_ensemble_td_read(const void* h, double* out) {
  const DzoEnsemble* e = (const DzoEnsemble*)h;
  std::memcpy(out, e->state.data(), e->state.size() * sizeof(double));
  return OK;
}
void dzo_ensemble_td_close(void* h) { delete (DzoEnsemble*)h; }

// ---- extended-precision Thomas ("truth" for the conditioning study of the 10^7-node bar, SURVEY §7 hard part 1).
// Same recurrences as solve_by_thomas, evaluated in __float128 (113-bit) or x87 long double (64-bit mantissa).
}  // extern "C"
template <class T>
static int thomas_ext(const double* lo, const double* di, const double* up, const double* b, size_t n, double* out) {
  if (n < 2) return WRONG_DIMS;
  std::vector<T> c(n - 1), d(n);
  c[0] = (T)up[0] / (T)di[0];
  d[0] = (T)b[0] / (T)di[0];
  for (size_t i = 1; i < n - 1; ++i) {
    T den = (T)di[i] - (T)lo[i] * c[i - 1];
    c[i] = (T)up[i] / den;
    d[i] = ((T)b[i] - (T)lo[i] * d[i - 1]) / den;
  }
  d[n - 1] = ((T)b[n - 1] - (T)lo[n - 1] * d[n - 2]) / ((T)di[n - 1] - (T)lo[n - 1] * c[n - 2]);
  T s = d[n - 1];
  out[n - 1] = (double)s;
  for (size_t i = n - 1; i-- > 0;) {
    s = d[i] - c[i] * s;
    out[i] = (double)s;
  }
  return OK;
}
extern "C" {
int dzo_thomas_f128(const double* lo, const double* di, const double* up, const double* b, size_t n, double* out) {
  return thomas_ext<__float128>(lo, di, up, b, n, out);
}
int dzo_thomas_f80(const double* lo, const double* di, const double* up, const double* b, size_t n, double* out) {
  return thomas_ext<long double>(lo, di, up, b, n, out);
}

// ---- "what the reference means to compute": its static diffusion system (time_independent.rs:91-182) evaluated in EXACT
// arithmetic (__float128: 113 bits, ~1e-34) on the reference's own f64 inputs — mesh coordinates, the rule's f64 nodes
// x_j = cos(theta_j) and weights w_j, j = 1..G-1 (:130-136), mu, b, the boundary values — and solved in __float128 without ever
// rounding a band entry to f64.  On a long bar the f64 rounding noise of the band entries, whoever assembles them, is the
// leading error of the solution; this is the yardstick that tells how far the reference's own f64 result and a
// differently rounded assembly each are from the system they both approximate.
// In exact arithmetic the three quadrature sums of interior row i (a = x_{i-1}, c = x_i, e = x_{i+1}; h1 = c-a, h2 = e-c, H = e-a)
// only depend on the rule through its moments (the integrands are first-degree polynomials of x_j on each piece):
//   prev = -mu T0 / (2 h1) - b (T0 + T1) / 4,    next = -mu T0 / (2 h2) + b (T0 - T1) / 4,
//   diag = (H/2) [ mu (WL/h1^2 + WR/h2^2) + b (H/2) ((WL + XL)/h1^2 - (WR - XR)/h2^2) ],
// T0 = sum w_j, T1 = sum w_j x_j; WL, XL (WR, XR) the same sums over the nodes whose image (a+e)/2 + (H/2) x_j lies left of c
// (not left of c): the piece phi_i is evaluated on (piecewise_polynomials_1degree.rs:133-148, strict <).
// brute != 0 evaluates the sums node by node instead (the reference's loop, every operation in __float128): the check of the
// closed form (tests/test_oracle.py).
}  // extern "C"
typedef __float128 q128;
static void intended_row(const double* mesh, size_t i, const std::vector<double>& xs, const std::vector<double>& ws,
                         const std::vector<q128>& pw, const std::vector<q128>& px, q128 mu, q128 b, int brute, q128* lo, q128* di,
                         q128* up) {
  const q128 a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  const q128 h1 = c - a, h2 = e - c, H = e - a;
  const size_t m = xs.size();
  if (brute) {
    q128 sp = 0, sn = 0, sd = 0;
    for (size_t j = 0; j < m; ++j) {
      const q128 x = xs[j], w = ws[j];
      const q128 tp = (a + c) / 2 + (h1 / 2) * x, tn = (c + e) / 2 + (h2 / 2) * x, ts = (a + e) / 2 + (H / 2) * x;
      // phi_i = (t - a)/h1 on [a, c), (e - t)/h2 on [c, e); phi_{i-1} = (c - t)/h1 on [a, c); phi_{i+1} = (t - c)/h2 on [c, e)
      sp += (mu * (1 / h1) * (-1 / h1) + b * (-1 / h1) * ((tp - a) / h1)) * (h1 / 2) * w;
      sn += (mu * (-1 / h2) * (1 / h2) + b * (1 / h2) * ((e - tn) / h2)) * (h2 / 2) * w;
      if (ts < c) sd += (mu * (1 / h1) * (1 / h1) + b * (1 / h1) * ((ts - a) / h1)) * (H / 2) * w;
      else sd += (mu * (-1 / h2) * (-1 / h2) + b * (-1 / h2) * ((e - ts) / h2)) * (H / 2) * w;
    }
    *lo = sp, *up = sn, *di = sd;
    return;
  }
  const q128 T0 = pw[m], T1 = px[m];
  // nodes decrease with j: the images left of the kink are j = s .. m-1, s = first j with (a+e)/2 + (H/2) x_j < c
  size_t l = 0, r = m;
  while (l < r) {
    const size_t mid = (l + r) / 2;
    if ((a + e) / 2 + (H / 2) * (q128)xs[mid] < c) r = mid;
    else l = mid + 1;
  }
  const q128 WR = pw[l], XR = px[l], WL = T0 - WR, XL = T1 - XR;
  *lo = -mu * T0 / (2 * h1) - b * (T0 + T1) / 4;
  *up = -mu * T0 / (2 * h2) + b * (T0 - T1) / 4;
  *di = (H / 2) * (mu * (WL / (h1 * h1) + WR / (h2 * h2)) + b * (H / 2) * ((WL + XL) / (h1 * h1) - (WR - XR) / (h2 * h2)));
}
// The mass matrix of the time-dependent solver (time_dependent.rs:180-190) the same way: second moments as well.
//   prev = (h1/8)(T0 - T2),  next = (h2/8)(T0 - T2),  diag = (H^3/8) [ (WL + 2 XL + QL)/h1^2 + (WR - 2 XR + QR)/h2^2 ],
// T2 = sum w_j x_j^2, QL / QR its parts left / right of the kink.  (S of that solver is minus the static system's rows, :226-233.)
static void intended_mass_row(const double* mesh, size_t i, const std::vector<double>& xs, const std::vector<double>& ws,
                              const std::vector<q128>& pw, const std::vector<q128>& px, const std::vector<q128>& pq, int brute,
                              q128* lo, q128* di, q128* up) {
  const q128 a = mesh[i - 1], c = mesh[i], e = mesh[i + 1];
  const q128 h1 = c - a, h2 = e - c, H = e - a;
  const size_t m = xs.size();
  if (brute) {
    q128 sp = 0, sn = 0, sd = 0;
    for (size_t j = 0; j < m; ++j) {
      const q128 x = xs[j], w = ws[j];
      const q128 tp = (a + c) / 2 + (h1 / 2) * x, tn = (c + e) / 2 + (h2 / 2) * x, ts = (a + e) / 2 + (H / 2) * x;
      sp += ((tp - a) / h1) * ((c - tp) / h1) * (h1 / 2) * w;
      sn += ((e - tn) / h2) * ((tn - c) / h2) * (h2 / 2) * w;
      const q128 v = ts < c ? (ts - a) / h1 : (e - ts) / h2;
      sd += (v * v) * (H / 2) * w;
    }
    *lo = sp, *up = sn, *di = sd;
    return;
  }
  const q128 T0 = pw[m], T1 = px[m], T2 = pq[m];
  size_t l = 0, r = m;
  while (l < r) {
    const size_t mid = (l + r) / 2;
    if ((a + e) / 2 + (H / 2) * (q128)xs[mid] < c) r = mid;
    else l = mid + 1;
  }
  const q128 WR = pw[l], XR = px[l], QR = pq[l], WL = T0 - WR, XL = T1 - XR, QL = T2 - QR;
  *lo = (h1 / 8) * (T0 - T2);
  *up = (h2 / 8) * (T0 - T2);
  *di = (H * H * H / 8) * ((WL + 2 * XL + QL) / (h1 * h1) + (WR - 2 * XR + QR) / (h2 * h2));
}
extern "C" {
// M and S of the time-dependent solver in exact arithmetic, rounded to f64 at the very end (identity rows at both ends, :236-239)
int dzo_td_intended_f128(const double* mesh, size_t n, double mu, double b, size_t G, int brute, int threads, double* mlo, double* mdi,
                         double* mup, double* slo, double* sdi, double* sup) {
  if (n < 3) return WRONG_DIMS;
  std::vector<double> xs, ws;
  int rc = quad_table(G, &xs, &ws);
  if (rc) return rc;
  const size_t m = xs.size();
  for (size_t j = 1; j < m; ++j)
    if (!(xs[j] < xs[j - 1])) return INTEGRATION;
  std::vector<q128> pw(m + 1), px(m + 1), pq(m + 1);
  pw[0] = 0, px[0] = 0, pq[0] = 0;
  for (size_t j = 0; j < m; ++j) {
    const q128 w = ws[j], x = xs[j];
    pw[j + 1] = pw[j] + w, px[j + 1] = px[j] + w * x, pq[j + 1] = pq[j] + w * x * x;
  }
  for (size_t i = 0; i < n; ++i) mlo[i] = mup[i] = slo[i] = sup[i] = 0.0, mdi[i] = sdi[i] = 1.0;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (long long i = 1; i < (long long)n - 1; ++i) {
    q128 l, d, u;
    intended_mass_row(mesh, (size_t)i, xs, ws, pw, px, pq, brute, &l, &d, &u);
    mlo[i] = (double)l, mdi[i] = (double)d, mup[i] = (double)u;
    intended_row(mesh, (size_t)i, xs, ws, pw, px, (q128)mu, (q128)b, brute, &l, &d, &u);
    slo[i] = (double)(-l), sdi[i] = (double)(-d), sup[i] = (double)(-u);
  }
  return OK;
}
// u (rounded to f64 at the very end) of the exact-arithmetic system; lo/di/up (optional, may be null): its bands rounded to f64
int dzo_ti_intended_f128(const double* mesh, size_t n, double mu, double b, double bc0, double bc1, size_t G, int brute, int threads,
                         double* u, double* lo, double* di, double* up) {
  if (n < 3) return WRONG_DIMS;
  std::vector<double> xs, ws;
  int rc = quad_table(G, &xs, &ws);
  if (rc) return rc;
  const size_t m = xs.size();
  for (size_t j = 1; j < m; ++j)
    if (!(xs[j] < xs[j - 1])) return INTEGRATION;  // the bisection wants decreasing nodes
  std::vector<q128> pw(m + 1), px(m + 1);  // prefix sums over j < s
  pw[0] = 0, px[0] = 0;
  for (size_t j = 0; j < m; ++j) pw[j + 1] = pw[j] + (q128)ws[j], px[j + 1] = px[j] + (q128)ws[j] * (q128)xs[j];
  std::vector<q128> L(n), D(n), U(n), F(n);
  L[0] = 0, D[0] = 1, U[0] = 0, F[0] = bc0;                  // :168-182: Dirichlet rows
  L[n - 1] = 0, D[n - 1] = 1, U[n - 1] = 0, F[n - 1] = bc1;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (long long i = 1; i < (long long)n - 1; ++i) {
    intended_row(mesh, (size_t)i, xs, ws, pw, px, (q128)mu, (q128)b, brute, &L[i], &D[i], &U[i]);
    F[i] = 0;
  }
  if (lo)
    for (size_t i = 0; i < n; ++i) lo[i] = (double)L[i], di[i] = (double)D[i], up[i] = (double)U[i];
  // solve_by_thomas's recurrences (matrix_solver/mod.rs:17-48), in place
  U[0] = U[0] / D[0], F[0] = F[0] / D[0];
  for (size_t i = 1; i < n - 1; ++i) {
    const q128 den = D[i] - L[i] * U[i - 1];
    U[i] = U[i] / den;
    F[i] = (F[i] - L[i] * F[i - 1]) / den;
  }
  F[n - 1] = (F[n - 1] - L[n - 1] * F[n - 2]) / (D[n - 1] - L[n - 1] * U[n - 2]);
  q128 s = F[n - 1];
  u[n - 1] = (double)s;
  for (size_t i = n - 1; i-- > 0;) {
    s = F[i] - U[i] * s;
    u[i] = (double)s;
  }
  return OK;
}

// ---- 2D ingestion ("next" row f3): mesh/mesh_builder.rs:342-500 (build_mesh_2d), :89-129 (obj_face_checker),
// :623-680 (merge_sort).  Boundary vertices = vertices of the edges that belong to exactly one triangle.
struct Mesh2D {
  std::vector<double> vertices;            // [a, b, 0, 0, 0, 1] per vertex
  std::vector<uint32_t> indices;           // 3 per triangle, 0-based
  std::vector<uint32_t> boundary_indices;  // ascending, unique
  double max_length = 0;
  float model_translation[3] = {0, 0, 0};
};
static bool rust_parse_u32(const std::string& s, uint32_t* out) {  // str::parse::<u32>: optional '+', digits only, no overflow
  size_t i = 0;
  if (i < s.size() && s[i] == '+') ++i;
  if (i >= s.size()) return false;
  unsigned long long v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (unsigned)(s[i] - '0');
    if (v > 0xffffffffULL) return false;
  }
  *out = (uint32_t)v;
  return true;
}
static int build_mesh_2d(const char* path, Mesh2D* m) {
  std::vector<std::string> lines;
  if (!read_lines(path, &lines)) return IO;
  std::vector<std::string> parts;
  std::set<std::string> keys[3];  // check_for_constant_coordinates :140-190
  for (const std::string& line : lines) {
    if (line.rfind("v ", 0) != 0) continue;
    split_single_space(line, &parts);
    if (parts.size() - 1 != 3) return MESH_PARSE;
    for (int a = 0; a < 3; ++a) {
      const std::string& c = parts[a + 1];
      double tmp;
      if (!rust_parse_f64(c, &tmp)) return MESH_PARSE;
      keys[a].insert((c.rfind("0.0", 0) == 0 || c.rfind("-0.0", 0) == 0) ? std::string("0.0") : c);
    }
  }
  int constant_coordinate;  // :355-363
  if (keys[0].size() == 1) constant_coordinate = 0;
  else if (keys[1].size() == 1) constant_coordinate = 1;
  else if (keys[2].size() == 1) constant_coordinate = 2;
  else return MESH_PARSE;
  double x_min = 0.0, y_min = 0.0, x_max = 0.0, y_max = 0.0;  // :366-371 — all start at 0.0
  std::map<std::pair<uint32_t, uint32_t>, size_t> boundary_edges;  // HashMap<[u32;2], usize> (iteration order is irrelevant: sorted later)
  auto count_edge = [&](uint32_t p, uint32_t q) {  // :417-446: (p,q), else (q,p), else insert (p,q)
    auto it = boundary_edges.find({p, q});
    if (it != boundary_edges.end()) { ++it->second; return; }
    it = boundary_edges.find({q, p});
    if (it != boundary_edges.end()) { ++it->second; return; }
    boundary_edges[{p, q}] = 1;
  };
  for (const std::string& line : lines) {
    if (line.rfind("v ", 0) == 0) {
      split_single_space(line, &parts);
      std::vector<double> coordinate;
      for (size_t t = 1; t < parts.size(); ++t) {
        double v;
        if (!rust_parse_f64(parts[t], &v)) return MESH_PARSE;
        coordinate.push_back(v);
      }
      if (coordinate.size() != 3) return MESH_PARSE;
      coordinate.erase(coordinate.begin() + constant_coordinate);  // :388-389
      coordinate.push_back(0.0);
      if (coordinate[0] < x_min) x_min = coordinate[0];  // :392-407
      if (coordinate[0] > x_max) x_max = coordinate[0];
      if (coordinate[1] < y_min) y_min = coordinate[1];
      if (coordinate[1] > y_max) y_max = coordinate[1];
      m->vertices.insert(m->vertices.end(), coordinate.begin(), coordinate.end());
      m->vertices.push_back(0.0), m->vertices.push_back(0.0), m->vertices.push_back(1.0);  // blue :411
    } else if (line.rfind("f ", 0) == 0) {
      split_single_space(line, &parts);  // obj_face_checker :89-129
      if (parts.size() - 1 != 3) return MESH_PARSE;
      uint32_t tri[3];
      for (int k = 0; k < 3; ++k) {
        std::vector<std::string> fp;
        size_t start = 0;
        const std::string& face = parts[k + 1];
        for (;;) {  // str::split("/")
          const size_t pos = face.find('/', start);
          if (pos == std::string::npos) { fp.push_back(face.substr(start)); break; }
          fp.push_back(face.substr(start, pos - start));
          start = pos + 1;
        }
        uint32_t v;
        if (!rust_parse_u32(fp[0], &v)) return MESH_PARSE;
        if (fp.size() - 1 != 2) return MESH_PARSE;
        tri[k] = v - 1;
      }
      count_edge(tri[0], tri[1]);
      count_edge(tri[0], tri[2]);
      count_edge(tri[2], tri[1]);
      m->indices.insert(m->indices.end(), tri, tri + 3);
    }
  }
  const double len_x = x_max - x_min, len_y = y_max - y_min;  // :455-461
  m->max_length = len_x > len_y ? len_x : len_y;
  m->model_translation[0] = (float)x_min + (float)m->max_length / 2.0f;  // :464-465
  m->model_translation[1] = (float)y_min + (float)m->max_length / 2.0f;
  m->model_translation[2] = 0.0f;
  std::set<uint32_t> b;  // :468-478 HashSet of the vertices of the edges counted once; merge_sort -> ascending
  for (const auto& kv : boundary_edges)
    if (kv.second == 1) b.insert(kv.first.first), b.insert(kv.first.second);
  if (b.empty()) return MESH_PARSE;  // the reference panics inside merge_sort on an empty vector
  m->boundary_indices.assign(b.begin(), b.end());
  return OK;
}

// ---- mesh ingestion.  Two-call protocol: n_out = number of nodes; buffers sized 12n / 6n-6.
int dzo_mesh_ingest_1d(const char* path, int has_mult, double mult, size_t* n_out, double* vertices, size_t vcap,
                       uint32_t* indices, size_t icap, double* max_length, double* prom_width, float* translation) {
  Mesh1D m;
  int rc = build_mesh_1d(path, has_mult != 0, mult, &m);
  if (rc) return rc;
  *n_out = m.vertices.size() / 12;
  if (vertices) {
    if (vcap < m.vertices.size() || icap < m.indices.size()) return WRONG_DIMS;
    std::memcpy(vertices, m.vertices.data(), m.vertices.size() * sizeof(double));
    std::memcpy(indices, m.indices.data(), m.indices.size() * sizeof(uint32_t));
    *max_length = m.max_length;
    *prom_width = m.prom_width;
    for (int k = 0; k < 3; ++k) translation[k] = m.model_translation[k];
  }
  return OK;
}
// Two-call protocol like dzo_mesh_ingest_1d: counts first (buffers NULL), then data.
int dzo_mesh_ingest_2d(const char* path, size_t* n_vertices, size_t* n_triangles, size_t* n_boundary, double* vertices,
                       uint32_t* indices, uint32_t* boundary, double* max_length, float* translation) {
  Mesh2D m;
  int rc = build_mesh_2d(path, &m);
  if (rc) return rc;
  *n_vertices = m.vertices.size() / 6, *n_triangles = m.indices.size() / 3, *n_boundary = m.boundary_indices.size();
  if (vertices) {
    std::memcpy(vertices, m.vertices.data(), m.vertices.size() * sizeof(double));
    std::memcpy(indices, m.indices.data(), m.indices.size() * sizeof(uint32_t));
    std::memcpy(boundary, m.boundary_indices.data(), m.boundary_indices.size() * sizeof(uint32_t));
    *max_length = m.max_length;
    for (int k = 0; k < 3; ++k) translation[k] = m.model_translation[k];
  }
  return OK;
}
// the same boundary detection on an in-memory triangle list (CPU baseline for large synthetic meshes)
int dzo_boundary_vertices(const uint32_t* tris, size_t n_triangles, uint32_t* out, size_t cap, size_t* n_out) {
  std::map<std::pair<uint32_t, uint32_t>, size_t> edges;
  auto count_edge = [&](uint32_t p, uint32_t q) {
    auto it = edges.find({p, q});
    if (it != edges.end()) { ++it->second; return; }
    it = edges.find({q, p});
    if (it != edges.end()) { ++it->second; return; }
    edges[{p, q}] = 1;
  };
  for (size_t t = 0; t < n_triangles; ++t) {
    count_edge(tris[3 * t], tris[3 * t + 1]);
    count_edge(tris[3 * t], tris[3 * t + 2]);
    count_edge(tris[3 * t + 2], tris[3 * t + 1]);
  }
  std::set<uint32_t> b;
  for (const auto& kv : edges)
    if (kv.second == 1) b.insert(kv.first.first), b.insert(kv.first.second);
  *n_out = b.size();
  if (b.size() > cap) return WRONG_DIMS;
  size_t k = 0;
  for (uint32_t v : b) out[k++] = v;
  return OK;
}
// mesh/mod.rs:54-68 — every 6th value of the first half
int dzo_filter_for_solving_1d(const double* vertices, size_t len, double* out) {
  const size_t vertices_len = len / 12;
  size_t k = 0;
  for (size_t idx = 0; idx < len; ++idx)
    if (idx % 6 == 0 && idx < vertices_len * 6) out[k++] = vertices[idx];
  return (int)k;
}
// mesh/mod.rs:73-90 — colour gradient written into slots 3 and 5 of both vertex copies ("next" row f1)
void dzo_update_gradient_1d(double* vertices, size_t len, const double* velocity_norm, size_t nv) {
  double sol_max = -INFINITY, sol_min = INFINITY;
  for (size_t i = 0; i < nv; ++i) {
    sol_max = std::fmax(sol_max, velocity_norm[i]);  // f64::max ignores NaN like fmax
    sol_min = std::fmin(sol_min, velocity_norm[i]);
  }
  for (size_t i = 0; i < len / 12; ++i) {
    const double norm_sol = (velocity_norm[i] - sol_min) / (sol_max - sol_min) * (dzo::PI / 2.);
    vertices[6 * i + 3] = std::sin(norm_sol);
    vertices[6 * i + 5] = std::cos(norm_sol);
    vertices[6 * i + 3 + len / 2] = std::sin(norm_sol);
    vertices[6 * i + 5 + len / 2] = std::cos(norm_sol);
  }
}

}  // extern "C"

// ============================================================== EulerSolver (solvers/euler/mod.rs:26-59)
// do_step with the caller's closure: values = [y^(n-1), ..., y, t], width 2..5 (FunctionArguments, :19-22).
extern "C" {
int dzo_euler_do_step(const double* values, size_t width, double step, double (*f)(const double* values, void* ctx), void* ctx,
                      double* next) {
  if (width < 2 || width > 5) return dzo::WRONG_DIMS;
  const double f_eval = f(values, ctx);                        // :38
  double value = values[0] + step * f_eval;                    // :43
  const double t_new = values[width - 1] + step;               // :44
  next[0] = value;                                             // :46
  for (size_t j = 1; j + 1 < width; ++j) {                     // :48-52
    const double new_val = values[j] + step * value;
    value = new_val;
    next[j] = new_val;
  }
  next[width - 1] = t_new;                                     // :53
  return dzo::OK;
}
// the closure  c[0] + c[1]*v[0] + ... + c[width]*v[width-1]  (left to right), stepped `steps` times on every trajectory;
// coeffs [n_traj][width + 1], values [n_traj][width] (updated in place)
static double dzo_affine_closure(const double* v, void* ctx) {
  const double* c = static_cast<const double*>(ctx);
  const size_t width = (size_t)c[-1];
  double f = c[0];
  for (size_t j = 0; j < width; ++j) f = f + c[j + 1] * v[j];
  return f;
}
int dzo_euler_affine_run(const double* coeffs, double* values, size_t n_traj, size_t width, double step, size_t steps) {
  if (width < 2 || width > 5) return dzo::WRONG_DIMS;
  for (size_t i = 0; i < n_traj; ++i) {
    double ctx[7];
    ctx[0] = (double)width;
    for (size_t j = 0; j <= width; ++j) ctx[1 + j] = coeffs[i * (width + 1) + j];
    double* v = values + i * width;
    for (size_t s = 0; s < steps; ++s) {
      double next[5];
      dzo_euler_do_step(v, width, step, dzo_affine_closure, ctx + 1, next);
      for (size_t j = 0; j < width; ++j) v[j] = next[j];
    }
  }
  return dzo::OK;
}
}  // extern "C"
