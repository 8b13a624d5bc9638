// gl_quadrature.cpp — see gl_quadrature.h.  Compiled with -ffp-contract=off (Rust never fuses a*b+c).
#include "gl_quadrature.h"

#include <cmath>

#include "gl_tables_generated.h"

namespace dzb {
namespace {

const double kPi = 3.14159265358979323846264338327950288;

// Polynomial in `t`, coefficients from the highest power down: ((c0*t + c1)*t + c2)...
template <int N>
inline double poly_desc(const double (&c)[N], double t) {
  double v = c[0];
  for (int i = 1; i < N; ++i) v = v * t + c[i];
  return v;
}
// Series a0 + t*(a1 + t*(a2 + ... + t*a_{N-1})), evaluated from the innermost bracket outwards.
template <int N>
inline double series_asc(const double (&a)[N], double t) {
  double v = a[N - 2] + a[N - 1] * t;
  for (int i = N - 3; i >= 0; --i) v = a[i] + t * v;
  return v;
}

// k-th positive zero of J0: tabulated below 20, McMahon's expansion above (gauss_legendre.rs:5684-5702)
double j0_zero(size_t k) {
  if (k < 20) return dzb_gl_tables::JZ[k - 1];
  static const double mcmahon[9] = {0.125,
                                    -0.807291666666666666666666666667e-1,
                                    0.246028645833333333333333333333,
                                    -1.82443876720610119047619047619,
                                    25.3364147973439050099206349206,
                                    -567.644412135183381139802038240,
                                    18690.4765282320653831636345064,
                                    -8.49353580299148769921876983660e5,
                                    5.09225462402226769498681286758e7};
  const double ak = kPi * ((double)k - 0.25);
  const double r = 1.0 / ak;
  return ak + r * series_asc(mcmahon, r * r);
}

// J1(j_{0,k})^2: tabulated below 21, asymptotic series above (gauss_legendre.rs:5713-5730)
double j1_squared_at_j0_zero(size_t k) {
  if (k < 21) return dzb_gl_tables::J1[k - 1];
  static const double tail[8] = {-0.303380429711290253026202643516e-3, 0.198924364245969295201137972743e-3,
                                 -0.228969902772111653038747229723e-3, 0.433710719130746277915572905025e-3,
                                 -0.123632349727175414724737657367e-2, 0.496101423268883102872271417616e-2,
                                 -0.266837393702323757700998557826e-1, 0.185395398206345628711318848386};
  const double t = 1.0 / ((double)k - 0.25);
  const double t2 = t * t;
  return t * (0.202642367284675542887758926420 + t2 * t2 * series_asc(tail, t2));
}

// Asymptotic node/weight for n >= 101 (gauss_legendre.rs:5743-5853)
void asymptotic_pair(size_t n, size_t k, double* theta_out, double* weight_out) {
  static const double node1[7] = {-1.29052996274280508473467968379e-12, 2.40724685864330121825976175184e-10,
                                  -3.13148654635992041468855740012e-8,  0.275573168962061235623801563453e-5,
                                  -0.148809523713909147898955880165e-3, 0.416666666665193394525296923981e-2,
                                  -0.416666666666662959639712457549e-1};
  static const double node2[7] = {2.20639421781871003734786884322e-9,  -7.53036771373769326811030753538e-8,
                                  0.161969259453836261731700382098e-5, -0.253300326008232025914059965302e-4,
                                  0.282116886057560434805998583817e-3, -0.209022248387852902722635654229e-2,
                                  0.815972221772932265640401128517e-2};
  static const double node3[7] = {-2.97058225375526229899781956673e-8, 5.55845330223796209655886325712e-7,
                                  -0.567797841356833081642185432056e-5, 0.418498100329504574443885193835e-4,
                                  -0.251395293283965914823026348764e-3, 0.128654198542845137196151147483e-2,
                                  -0.416012165620204364833694266818e-2};
  static const double wgt1[10] = {-2.20902861044616638398573427475e-14, 2.30365726860377376873232578871e-12,
                                  -1.75257700735423807659851042318e-10, 1.03756066927916795821098009353e-8,
                                  -4.63968647553221331251529631098e-7,  0.149644593625028648361395938176e-4,
                                  -0.326278659594412170300449074873e-3, 0.436507936507598105249726413120e-2,
                                  -0.305555555555553028279487898503e-1, 0.833333333333333302184063103900e-1};
  static const double wgt2[9] = {3.63117412152654783455929483029e-12, 7.67643545069893130779501844323e-11,
                                 -7.12912857233642220650643150625e-9, 2.11483880685947151466370130277e-7,
                                 -0.381817918680045468483009307090e-5, 0.465969530694968391417927388162e-4,
                                 -0.407297185611335764191683161117e-3, 0.268959435694729660779984493795e-2,
                                 -0.111111111111214923138249347172e-1};
  static const double wgt3[9] = {2.01826791256703301806643264922e-9,  -4.38647122520206649251063212545e-8,
                                 5.08898347288671653137451093208e-7,  -0.397933316519135275712977531366e-5,
                                 0.200559326396458326778521795392e-4, -0.422888059282921161626339411388e-4,
                                 -0.105646050254076140548678457002e-3, -0.947969308958577323145923317955e-4,
                                 0.656966489926484797412985260842e-2};
  const double w = 1.0 / ((double)n + 0.5);
  const double nu = j0_zero(k);
  double theta = w * nu;
  const double x = theta * theta;
  const double bj = j1_squared_at_j0_zero(k);
  const double f1 = poly_desc(node1, x), f2 = poly_desc(node2, x), f3 = poly_desc(node3, x);
  const double g1 = poly_desc(wgt1, x), g2 = poly_desc(wgt2, x), g3 = poly_desc(wgt3, x);
  const double nu_over_sin = nu / std::sin(theta);
  const double b_nu_over_sin = bj * nu_over_sin;
  const double w_inv_sinc = w * w * nu_over_sin;
  const double wis2 = w_inv_sinc * w_inv_sinc;
  theta = w * (nu + theta * w_inv_sinc * (f1 + wis2 * (f2 + wis2 * f3)));
  const double deno = b_nu_over_sin + b_nu_over_sin * wis2 * (g1 + wis2 * (g2 + wis2 * g3));
  *theta_out = theta;
  *weight_out = (2.0 * w) / deno;
}

// Tabulated pair for l <= 100, zero-based k (gauss_legendre.rs:5864-5896).  Row (l/2 - 1) resp. ((l-1)/2 - 1)
// of the even/odd tables has l/2 entries ordered by decreasing theta.
void tabulated_pair(size_t l, size_t k, double* theta, double* weight) {
  using namespace dzb_gl_tables;
  const bool odd = (l & 1) != 0;
  const size_t half = odd ? (l - 1) / 2 : l / 2;
  if (odd && k == half) {
    *theta = kPi / 2.0;
    *weight = 2.0 / (CL[l] * CL[l]);
    return;
  }
  const double* th = odd ? ODD_THETA + ODD_OFF[half - 1] : EVEN_THETA + EVEN_OFF[half - 1];
  const double* ww = odd ? ODD_W + ODD_OFF[half - 1] : EVEN_W + EVEN_OFF[half - 1];
  if (k < half) {
    *theta = th[half - k - 1];
    *weight = ww[half - k - 1];
  } else {
    const size_t idx = odd ? k - half - 1 : k - half;
    *theta = kPi - th[idx];
    *weight = ww[idx];
  }
}

}  // namespace

bool gl_quad_pair(size_t n, size_t k, double* theta, double* weight) {
  if (k == 0 || !(k < n)) return false;  // "Misuse of quad_pair function: k should be smaller than n"
  if (n < 101) {
    tabulated_pair(n, k - 1, theta, weight);
  } else if (2 * k - 1 > n) {
    asymptotic_pair(n, n - k + 1, theta, weight);
    *theta = kPi - *theta;
  } else {
    asymptotic_pair(n, k, theta, weight);
  }
  return true;
}

bool gl_build_rule(size_t G, QuadRule* r) {
  r->G = G;
  r->m = G >= 2 ? G - 1 : 0;
  const size_t m = r->m;
  r->x.assign(m, 0.0);
  r->w.assign(m, 0.0);
  for (size_t k = 1; k <= m; ++k) {
    double th, wt;
    if (!gl_quad_pair(G, k, &th, &wt)) return false;
    r->x[k - 1] = std::cos(th);  // `theta.cos()` (time_independent.rs:135): libm, on the host
    r->w[k - 1] = wt;
  }
  for (std::vector<double>* v : {&r->u0, &r->u1, &r->u2, &r->v0, &r->v1, &r->v2}) v->assign(m + 1, 0.0);
  long double a0 = 0, a1 = 0, a2 = 0;
  for (size_t j = m; j-- > 0;) {  // suffix sums of w (1+x)^p
    const long double ww = r->w[j], t = 1.0L + (long double)r->x[j];
    a0 += ww, a1 += ww * t, a2 += ww * t * t;
    r->u0[j] = (double)a0, r->u1[j] = (double)a1, r->u2[j] = (double)a2;
  }
  a0 = a1 = a2 = 0;
  long double qq = 0;
  for (size_t j = 0; j < m; ++j) {  // prefix sums of w (1-x)^p
    const long double ww = r->w[j], xx = r->x[j], t = 1.0L - xx;
    a0 += ww, a1 += ww * t, a2 += ww * t * t;
    qq += ww * (1.0L - xx * xx);
    r->v0[j + 1] = (double)a0, r->v1[j + 1] = (double)a1, r->v2[j + 1] = (double)a2;
  }
  r->t0 = m ? r->u0[0] : 0.0;
  r->t1p = m ? r->u1[0] : 0.0;
  r->t1m = r->v1[m];
  r->q = (double)qq;
  return true;
}

}  // namespace dzb
