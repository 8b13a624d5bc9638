// bigsys_kernels.cuh — one long tridiagonal system (a single bar with n >> 1024 nodes).
// v0: correctness scaffold that runs the faithful recurrences on one thread (k_thomas_factor /
// k_thomas_solve_faithful / k_td_step_faithful).  Replaced by the tiled scan solver — the interface
// below is what dzb_api.cu programs against.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "tridiag_kernels.cuh"

namespace dzb {

struct BigSys {
  size_t n = 0;
  double* c = nullptr;
  double* den = nullptr;
};

inline std::string& bigsys_error_ref() {
  static thread_local std::string e;
  return e;
}
inline const char* bigsys_error() { return bigsys_error_ref().c_str(); }
inline int bigsys_fail(cudaError_t e, const char* what) {
  bigsys_error_ref() = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}

inline void bigsys_free(BigSys* s) {
  if (s->c) cudaFree(s->c);
  if (s->den) cudaFree(s->den);
  s->c = s->den = nullptr;
  s->n = 0;
}

inline int bigsys_factor(BigSys* s, const double* lo, const double* di, const double* up, size_t n, cudaStream_t st,
                         unsigned long long* launches) {
  s->n = n;
  cudaError_t e;
  if ((e = cudaMalloc((void**)&s->c, n * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
  if ((e = cudaMalloc((void**)&s->den, n * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
  k_thomas_factor<<<1, 32, 0, st>>>(lo, di, up, s->c, s->den, (long long)n, 1);
  ++*launches;
  if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_thomas_factor");
  return 0;
}

inline int bigsys_solve(BigSys* s, const double* lo, const double* di, const double* up, const double* rhs, double* out,
                        int refine_steps, cudaStream_t st, unsigned long long* launches) {
  (void)di, (void)up, (void)refine_steps;
  k_thomas_solve_faithful<<<1, 32, 0, st>>>(lo, s->c, s->den, rhs, out, (long long)s->n, 1);
  ++*launches;
  cudaError_t e;
  if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_thomas_solve_faithful");
  return 0;
}

inline int bigsys_td_step(BigSys* s, const double* u, double* u_new, const double* mlo, const double* mdi, const double* mup,
                          const double* slo, const double* sdi, const double* sup, double dt, double bc0, double bc1,
                          cudaStream_t st, unsigned long long* launches) {
  k_td_step_faithful<<<1, 32, 0, st>>>(u, u_new, mlo, mdi, mup, slo, sdi, sup, s->c, s->den, (long long)s->n, 1, dt, bc0, bc1);
  ++*launches;
  cudaError_t e;
  if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_td_step_faithful");
  return 0;
}

}  // namespace dzb
