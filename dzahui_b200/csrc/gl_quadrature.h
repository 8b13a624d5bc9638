// gl_quadrature.h — host-side Gauss-Legendre rule of the reference (Bogaert 2014, iteration-free), product copy.
// Follows quadrature/gauss_legendre.rs:5684-5927 of the reference operation for operation so that the table
// uploaded to the GPU carries the very same f64 (cos(theta_k), w_k) the Rust integration loops consume.
#pragma once
#include <cstddef>
#include <vector>

namespace dzb {

// quad_pair(n, k), 1 <= k < n  (gauss_legendre.rs:5908-5927).  Returns false for the rejected k >= n (and k == 0).
bool gl_quad_pair(size_t n, size_t k, double* theta, double* weight);

// Everything the device kernels need about the G-point rule restricted to the nodes k = 1..G-1 that the
// reference's `for j in 1..gauss_step` loops visit (time_independent.rs:132, time_dependent.rs:166, dim1.rs:156,210).
struct QuadRule {
  size_t G = 0;
  size_t m = 0;            // number of used nodes = G - 1 (0 when G < 2)
  std::vector<double> x;   // x[k-1] = cos(theta_k), strictly decreasing in k
  std::vector<double> w;   // w[k-1]
  // Moments for the closed-form (FAST) assembly, accumulated in long double and rounded once.  With 0-based node
  // index j (node k = j+1) and a split position s in [0, m]:
  //   U_p[s] = sum_{j >= s} w_j (1 + x_j)^p     (the nodes that fall LEFT of the kink: x decreasing in j)
  //   V_p[s] = sum_{j <  s} w_j (1 - x_j)^p     (the nodes that fall RIGHT of the kink)
  std::vector<double> u0, u1, u2, v0, v1, v2;  // each m + 1 entries
  double t0 = 0, t1p = 0, t1m = 0, q = 0;      // sum w, sum w(1+x), sum w(1-x), sum w(1-x^2)
};
bool gl_build_rule(size_t G, QuadRule* rule);

}  // namespace dzb
