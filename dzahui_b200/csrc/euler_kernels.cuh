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

// ---- the derivative function as a postfix program ------------------------------------------------------------------------
// The reference's closure is any `Fn(&[f64; N]) -> f64` (mod.rs:24-31).  For derivative functions that are not affine the
// device takes the expression itself, compiled by the host into a short postfix program over the OLD values, per-trajectory
// parameters and constants (dzb_euler_create_program).  IEEE double, one rounding per operation in program order (the
// translation unit is built with -fmad=false): +, -, *, /, neg, abs, sqrt give the bits the Rust closure gives; sin, cos, exp,
// log come from CUDA's libm instead of the host's (a last-place difference at most).
enum { EOP_CONST = 0, EOP_VALUE = 1, EOP_PARAM = 2, EOP_ADD = 3, EOP_SUB = 4, EOP_MUL = 5, EOP_DIV = 6, EOP_NEG = 7, EOP_SQRT = 8,
       EOP_ABS = 9, EOP_SIN = 10, EOP_COS = 11, EOP_EXP = 12, EOP_LOG = 13, EOP_COUNT = 14 };
constexpr int EULER_MAX_OPS = 96, EULER_MAX_STACK = 16;
struct EulerOp {
  int opcode, arg;
  double value;
};
struct EulerProgram {
  const EulerOp* ops;  // device
  int n_ops, n_params;
  const double* params;  // SoA [n_params][T]
};

template <int W>
__global__ void __launch_bounds__(128) k_euler_program(double* __restrict__ v, const EulerProgram prog, long long T, double step, long long k) {
  __shared__ EulerOp ops[EULER_MAX_OPS];
  for (int j = threadIdx.x; j < prog.n_ops; j += blockDim.x) ops[j] = prog.ops[j];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  double x[W];
#pragma unroll
  for (int j = 0; j < W; ++j) x[j] = v[(long long)j * T + i];
  for (long long s = 0; s < k; ++s) {
    double st[EULER_MAX_STACK];
    int sp = 0;  // the host has checked the program: the stack never under- or overflows and ends with one entry
    for (int q = 0; q < prog.n_ops; ++q) {
      const EulerOp op = ops[q];
      switch (op.opcode) {
        case EOP_CONST: st[sp++] = op.value; break;
        case EOP_VALUE: {
          double val = x[0];
#pragma unroll
          for (int j = 1; j < W; ++j) val = op.arg == j ? x[j] : val;
          st[sp++] = val;
          break;
        }
        case EOP_PARAM: st[sp++] = prog.params[(long long)op.arg * T + i]; break;
        case EOP_ADD: --sp, st[sp - 1] = st[sp - 1] + st[sp]; break;
        case EOP_SUB: --sp, st[sp - 1] = st[sp - 1] - st[sp]; break;
        case EOP_MUL: --sp, st[sp - 1] = st[sp - 1] * st[sp]; break;
        case EOP_DIV: --sp, st[sp - 1] = st[sp - 1] / st[sp]; break;
        case EOP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case EOP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
        case EOP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
        case EOP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
        case EOP_COS: st[sp - 1] = cos(st[sp - 1]); break;
        case EOP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
        default: st[sp - 1] = log(st[sp - 1]); break;
      }
    }
    const double f = st[0];                               // the closure, on the OLD values (mod.rs:38)
    const double t_new = x[W - 1] + step;                 // :44
    double value = x[0] + step * f;                       // :43
    x[0] = value;
#pragma unroll
    for (int j = 1; j < W - 1; ++j) {                     // :48-52
      value = x[j] + step * value;
      x[j] = value;
    }
    x[W - 1] = t_new;
  }
#pragma unroll
  for (int j = 0; j < W; ++j) v[(long long)j * T + i] = x[j];
}

}  // namespace dzb
