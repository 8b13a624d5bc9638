// euler_kernels.cuh — EulerSolver::do_step (reference solvers/euler/mod.rs:37-59) for an ensemble of trajectories.
//
// The reference's stepper takes values = [y^(n-1), ..., y', y, t] (2 to 5 entries, mod.rs:19-22), evaluates the user's
// closure f(&values) once and cascades:  next[0] = values[0] + step f;  next[j] = values[j] + step next[j-1]  for the
// middle entries;  t + step.  A host closure cannot run on the device, so the device version covers derivative functions
// that are AFFINE in the state,  f(values) = c[0] + c[1] values[0] + ... + c[W] values[W-1]  (evaluated left to right, one
// rounding per operation: what the closure `c0 + c1*v[0] + ...` compiles to) — which includes every closure of the
// reference's own integration tests (tests/euler.rs: 1.0, -9.81, -lambda * val[0]).  One thread per trajectory, state and
// coefficients in registers for all k steps of a call; SoA [W][T] so that the loads and stores are coalesced.
#pragma once
#include <cuda_runtime.h>

namespace dzb {

template <int W>
__global__ void __launch_bounds__(256) k_euler_steps(double* __restrict__ v, const double* __restrict__ c, long long T, double step,
                                                     long long k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  double x[W], a[W + 1];
#pragma unroll
  for (int j = 0; j < W; ++j) x[j] = v[(long long)j * T + i];
#pragma unroll
  for (int j = 0; j <= W; ++j) a[j] = c[(long long)j * T + i];
  for (long long s = 0; s < k; ++s) {
    double f = a[0];
#pragma unroll
    for (int j = 0; j < W; ++j) f = f + a[j + 1] * x[j];  // the closure, on the OLD values (mod.rs:38)
    const double t_new = x[W - 1] + step;                 // :44
    double value = x[0] + step * f;                       // :43
    x[0] = value;
#pragma unroll
    for (int j = 1; j < W - 1; ++j) {                     // :48-52: each entry integrates the one just updated
      value = x[j] + step * value;
      x[j] = value;
    }
    x[W - 1] = t_new;
  }
#pragma unroll
  for (int j = 0; j < W; ++j) v[(long long)j * T + i] = x[j];
}

}  // namespace dzb
