// bigsys_kernels.cuh — one long tridiagonal system (a single bar with n >> 1024 nodes): Thomas / PCR-style hybrid.
//
// The reference solves K u = b with the serial Thomas recurrence (matrix_solver/mod.rs:17-48), which cannot use a GPU
// at n = 10^7.  Here every thread runs a *modified Thomas* sweep over R = 8 consecutive rows that eliminates the six
// interior unknowns and leaves two equations coupling the chunk's first/last unknown with the neighbouring chunks:
//     s-row:  lo_s u_{s-1} + (di_s - up_s a_1) u_s - up_s c_1 u_e           = f_s - up_s d_1
//     e-row:               a_7 u_s           +      u_e + c_7 u_{e+1}       = d_7
// which is again a tridiagonal system, of a quarter of the size, in the order (s_0, e_0, s_1, e_1, ...).  The interior
// values follow without division, u_j = d_j - a_j u_s - c_j u_e.  The reduction is applied recursively:
//     level 0   registers      2048 rows per CTA (256 threads x 8 rows, 128-bit loads)
//     level 1-4 shared memory  512 -> 128 -> 32 -> 8 -> 2 rows, chunk-interleaved SoA (conflict-free)
//     grid      k_tri_tile<REDUCE> writes 2 rows per tile; k_tri_mid (one CTA, shared memory only) solves that system
//               (recursing through k_tri_tile again while it exceeds 2048 rows); k_tri_tile<SOLVE> redoes the up-sweep on chip and walks back
//               down.  HBM traffic: 32 B/row (reduce) + 40 B/row (solve) for materialised bands.
// Pivoting: none, like the reference; chunk-local pivots are the diagonals of the diffusion / mass matrices
// (M-matrices), not the cancelling diagonal of the Stokes matrix, which therefore stays on the serial kernel.
// Accuracy: a partitioned elimination loses ~2 digits against serial Thomas on the n = 10^7 stiffness matrix (see
// DESIGN.md), so static solves refine once with a double-double residual (k_residual_dd), which lands closer to the
// exact solution of the reference's system than the reference's own f64 Thomas does.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "tridiag_kernels.cuh"

namespace dzb {

constexpr int BS_R = 8;                       // rows per chunk at every level
constexpr int BS_THREADS = 256;               // tile kernel
constexpr int BS_TILE = BS_R * BS_THREADS;    // 2048 rows per CTA
constexpr int BS_MID_CAP = 2048;              // rows the single-CTA kernel k_tri_mid solves directly

// Shared-memory level: 8*P rows, SoA, row r of the level at (r % 8) * P + r / 8.
struct LevelPtr {
  double *a, *b, *c, *d;
};
__device__ __forceinline__ int lvl_pos(int r, int P) { return (r & 7) * P + (r >> 3); }

// One chunk (thread t of P) of a shared-memory level: modified Thomas in place, reduced rows -> next level (PN chunks;
// PN == 0: the next level is the final 2-row system stored at index 0/1).
__device__ __forceinline__ void level_reduce(const LevelPtr& lv, int P, const LevelPtr& nx, int PN, int t) {
  const double L0 = lv.a[t], D0 = lv.b[t], U0 = lv.c[t], F0 = lv.d[t];
  int idx = P + t;
  double r = __drcp_rn(lv.b[idx]);
  double aj = lv.a[idx] * r, cj = lv.c[idx] * r, dj = lv.d[idx] * r;
  lv.a[idx] = aj, lv.c[idx] = cj, lv.d[idx] = dj;
#pragma unroll
  for (int j = 2; j < BS_R; ++j) {
    idx += P;
    const double l = lv.a[idx];
    r = __drcp_rn(fma(-l, cj, lv.b[idx]));
    aj = -l * aj * r;
    dj = fma(-l, dj, lv.d[idx]) * r;
    cj = lv.c[idx] * r;
    lv.a[idx] = aj, lv.c[idx] = cj, lv.d[idx] = dj;
  }
  const double ea = aj, ec = cj, ed = dj;
  idx = 6 * P + t;
  double a1 = lv.a[idx], c1 = lv.c[idx], d1 = lv.d[idx];
#pragma unroll
  for (int j = BS_R - 3; j >= 1; --j) {
    idx -= P;
    const double cc = lv.c[idx];
    d1 = fma(-cc, d1, lv.d[idx]);
    a1 = fma(-cc, a1, lv.a[idx]);
    c1 = -cc * c1;
    lv.a[idx] = a1, lv.c[idx] = c1, lv.d[idx] = d1;
  }
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  nx.a[p0] = L0, nx.b[p0] = fma(-U0, a1, D0), nx.c[p0] = -U0 * c1, nx.d[p0] = fma(-U0, d1, F0);
  nx.a[p1] = ea, nx.b[p1] = 1.0, nx.c[p1] = ec, nx.d[p1] = ed;
}
// Down-sweep of the same chunk: the next level's d[] holds the solved u_s / u_e; this level's d[] receives u.
__device__ __forceinline__ void level_backsolve(const LevelPtr& lv, int P, const LevelPtr& nx, int PN, int t) {
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  const double us = nx.d[p0], ue = nx.d[p1];
  lv.d[t] = us;
  lv.d[7 * P + t] = ue;
#pragma unroll
  for (int j = 1; j < BS_R - 1; ++j) {
    const int idx = j * P + t;
    lv.d[idx] = fma(-lv.c[idx], ue, fma(-lv.a[idx], us, lv.d[idx]));
  }
}

// Shared-memory pyramid below a level-1 system of 8*P1 rows; P1 is a power of 4 (64 for tiles, 256 for k_tri_mid).
template <int P>
struct PyramidRows {
  static constexpr int value = 8 * P + PyramidRows<P / 4>::value;
};
template <>
struct PyramidRows<0> {
  static constexpr int value = 2;
};
template <int P1>
struct Pyramid {
  static constexpr int ROWS = PyramidRows<P1>::value;
  double a[ROWS], b[ROWS], c[ROWS], d[ROWS];
  __device__ __forceinline__ LevelPtr level(int off) { return {a + off, b + off, c + off, d + off}; }
  // up-sweep from level 1 to the final 2 rows (all threads of the CTA must call; needs >= P1 threads)
  __device__ __forceinline__ void reduce(int t) {
    int off = 0;
#pragma unroll
    for (int p = P1; p >= 1; p /= 4) {
      __syncthreads();
      if (t < p) level_reduce(level(off), p, level(off + 8 * p), p / 4, t);
      off += 8 * p;
    }
    __syncthreads();
  }
  __device__ __forceinline__ LevelPtr top() { return level(ROWS - 2); }
  // down-sweep; top().d[0..1] must hold the solved pair
  __device__ __forceinline__ void backsolve(int t) {
    int off = ROWS - 2;
#pragma unroll
    for (int p = 1; p <= P1; p *= 4) {
      off -= 8 * p;
      __syncthreads();
      if (t < p) level_backsolve(level(off), p, level(off + 8 * p), p / 4, t);
    }
    __syncthreads();
  }
};

// ---- tile kernel --------------------------------------------------------------------------------------------------
// SOLVE == false: up-sweep, the tile's two reduced rows -> red_{lo,di,up,f}[2 tile + {0,1}].
// SOLVE == true : up-sweep again, take (u_s, u_e) of the tile from ured[2 tile + {0,1}], down-sweep, write out[]
//                 (ACCUM: out[] += value, the iterative-refinement update).
template <bool SOLVE, bool ACCUM>
__global__ void __launch_bounds__(BS_THREADS, 3) k_tri_tile(const double* __restrict__ lo, const double* __restrict__ di,
                                                         const double* __restrict__ up, const double* __restrict__ f,
                                                         long long m, double* __restrict__ red_lo,
                                                         double* __restrict__ red_di, double* __restrict__ red_up,
                                                         double* __restrict__ red_f, const double* __restrict__ ured,
                                                         double* __restrict__ out) {
  __shared__ Pyramid<BS_THREADS / 4> pyr;  // level 1: 512 rows = 64 chunks
  const int t = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * BS_TILE + (long long)t * BS_R;
  double L[BS_R], D[BS_R], U[BS_R], F[BS_R];
  if (g0 + BS_R <= m) {
    const double2* pl = reinterpret_cast<const double2*>(lo + g0);
    const double2* pd = reinterpret_cast<const double2*>(di + g0);
    const double2* pu = reinterpret_cast<const double2*>(up + g0);
    const double2* pf = reinterpret_cast<const double2*>(f + g0);
#pragma unroll
    for (int j = 0; j < BS_R / 2; ++j) {
      const double2 vl = __ldcs(pl + j), vd = __ldcs(pd + j), vu = __ldcs(pu + j), vf = __ldcs(pf + j);
      L[2 * j] = vl.x, L[2 * j + 1] = vl.y, D[2 * j] = vd.x, D[2 * j + 1] = vd.y;
      U[2 * j] = vu.x, U[2 * j + 1] = vu.y, F[2 * j] = vf.x, F[2 * j + 1] = vf.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const long long g = g0 + j;
      const bool in = g < m;
      L[j] = in ? lo[g] : 0.0, D[j] = in ? di[g] : 1.0, U[j] = in ? up[g] : 0.0, F[j] = in ? f[g] : 0.0;
    }
  }
  if (g0 == 0) L[0] = 0.0;
  if (m - 1 >= g0 && m - 1 < g0 + BS_R) {
#pragma unroll
    for (int j = 0; j < BS_R; ++j)
      if (g0 + j == m - 1) U[j] = 0.0;
  }
  // level 0 in registers: (L, U, F)[1..7] become (a, c, d)
  const double L0 = L[0], D0 = D[0], U0 = U[0], F0 = F[0];
  {
    double r = __drcp_rn(D[1]);
    L[1] *= r, U[1] *= r, F[1] *= r;
#pragma unroll
    for (int j = 2; j < BS_R; ++j) {
      const double l = L[j];
      r = __drcp_rn(fma(-l, U[j - 1], D[j]));
      L[j] = -l * L[j - 1] * r;
      F[j] = fma(-l, F[j - 1], F[j]) * r;
      U[j] *= r;
    }
  }
  const double ea = L[BS_R - 1], ec = U[BS_R - 1], ed = F[BS_R - 1];
#pragma unroll
  for (int j = BS_R - 3; j >= 1; --j) {
    const double cc = U[j];
    F[j] = fma(-cc, F[j + 1], F[j]);
    L[j] = fma(-cc, L[j + 1], L[j]);
    U[j] = -cc * U[j + 1];
  }
  constexpr int P1 = BS_THREADS / 4;
  {
    LevelPtr l1 = pyr.level(0);
    const int p0 = lvl_pos(2 * t, P1), p1 = lvl_pos(2 * t + 1, P1);
    l1.a[p0] = L0, l1.b[p0] = fma(-U0, L[1], D0), l1.c[p0] = -U0 * U[1], l1.d[p0] = fma(-U0, F[1], F0);
    l1.a[p1] = ea, l1.b[p1] = 1.0, l1.c[p1] = ec, l1.d[p1] = ed;
  }
  pyr.reduce(t);
  LevelPtr top = pyr.top();
  if (!SOLVE) {
    if (t < 2) {
      const long long o = 2LL * blockIdx.x + t;
      red_lo[o] = top.a[t], red_di[o] = top.b[t], red_up[o] = top.c[t], red_f[o] = top.d[t];
    }
    return;
  }
  if (t < 2) top.d[t] = ured[2LL * blockIdx.x + t];
  pyr.backsolve(t);
  {
    LevelPtr l1 = pyr.level(0);
    const double us = l1.d[lvl_pos(2 * t, P1)], ue = l1.d[lvl_pos(2 * t + 1, P1)];
    double o[BS_R];
    o[0] = us, o[BS_R - 1] = ue;
#pragma unroll
    for (int j = 1; j < BS_R - 1; ++j) o[j] = fma(-U[j], ue, fma(-L[j], us, F[j]));
    if (g0 + BS_R <= m) {
      double2* po = reinterpret_cast<double2*>(out + g0);
#pragma unroll
      for (int j = 0; j < BS_R / 2; ++j) {
        double2 v = make_double2(o[2 * j], o[2 * j + 1]);
        if (ACCUM) {
          const double2 w = po[j];
          v.x += w.x, v.y += w.y;
        }
        po[j] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < BS_R; ++j)
        if (g0 + j < m) out[g0 + j] = ACCUM ? out[g0 + j] + o[j] : o[j];
    }
  }
}

// ---- single-CTA kernel: solves an m-row system (m <= 8 P1 <= 2048) entirely in shared memory ------------------------
// The rows are loaded coalesced straight into level 1 of the pyramid (identity rows pad up to 8 P1); inputs are not
// modified.  P1 in {4, 16, 64, 256}: 32 / 128 / 512 / 2048 rows.
template <int P1, bool ACCUM>
__global__ void __launch_bounds__(BS_THREADS) k_tri_mid(const double* __restrict__ lo, const double* __restrict__ di,
                                                        const double* __restrict__ up, const double* __restrict__ f, int m,
                                                        double* __restrict__ out) {
  extern __shared__ double4 mid_smem[];
  Pyramid<P1>& pyr = *reinterpret_cast<Pyramid<P1>*>(mid_smem);
  const int t = threadIdx.x;
  LevelPtr l1 = pyr.level(0);
  for (int r = t; r < 8 * P1; r += BS_THREADS) {
    const int p = lvl_pos(r, P1);
    const bool in = r < m;
    l1.a[p] = (in && r > 0) ? lo[r] : 0.0;
    l1.b[p] = in ? di[r] : 1.0;
    l1.c[p] = (in && r < m - 1) ? up[r] : 0.0;
    l1.d[p] = in ? f[r] : 0.0;
  }
  pyr.reduce(t);
  LevelPtr top = pyr.top();
  if (t == 0) {  // closed 2 x 2 system: [b0 c0; a1 b1] (u_s, u_e) = (f0, f1)
    const double b0 = top.b[0], c0 = top.c[0], a1 = top.a[1], b1 = top.b[1], f0 = top.d[0], f1 = top.d[1];
    const double det = fma(b0, b1, -c0 * a1);
    top.d[0] = fma(f0, b1, -c0 * f1) / det;
    top.d[1] = fma(b0, f1, -a1 * f0) / det;
  }
  pyr.backsolve(t);
  for (int r = t; r < m; r += BS_THREADS) {
    const double v = l1.d[lvl_pos(r, P1)];
    out[r] = ACCUM ? out[r] + v : v;
  }
}

// ---- iterative refinement: r = f - K u with a double-double accumulator -------------------------------------------
__device__ __forceinline__ void dd_sub_prod(double& s, double& comp, double x, double y) {  // (s, comp) -= x * y
  const double p = x * y;
  const double pe = fma(x, y, -p);  // x*y = p + pe exactly
  const double tt = s - p;
  const double bb = tt - s;
  const double se = (s - (tt - bb)) + (-p - bb);  // s - p = tt + se exactly
  s = tt;
  comp += se - pe;
}
__global__ void __launch_bounds__(256) k_residual_dd(const double* __restrict__ lo, const double* __restrict__ di,
                                                     const double* __restrict__ up, const double* __restrict__ f,
                                                     const double* __restrict__ u, double* __restrict__ r, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = f[i], comp = 0.0;
    if (i > 0) dd_sub_prod(s, comp, lo[i], u[i - 1]);
    dd_sub_prod(s, comp, di[i], u[i]);
    if (i + 1 < n) dd_sub_prod(s, comp, up[i], u[i + 1]);
    r[i] = s + comp;
  }
}

// ---- explicit-Euler right-hand side of one long bar (utils.rs:41-62, :16-30): b = dt (S u) + (M u) ---------------
__global__ void __launch_bounds__(256) k_td_rhs(const double* __restrict__ u, const double* __restrict__ mlo,
                                                const double* __restrict__ mdi, const double* __restrict__ mup,
                                                const double* __restrict__ slo, const double* __restrict__ sdi,
                                                const double* __restrict__ sup, double dt, double* __restrict__ b,
                                                long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double um = i > 0 ? u[i - 1] : 0.0, uc = u[i], up1 = i + 1 < n ? u[i + 1] : 0.0;
    double s, mm;
    if (i == 0) s = sdi[i] * uc + sup[i] * up1, mm = mdi[i] * uc + mup[i] * up1;
    else if (i == n - 1) s = slo[i] * um + sdi[i] * uc, mm = mlo[i] * um + mdi[i] * uc;
    else s = slo[i] * um + sdi[i] * uc + sup[i] * up1, mm = mlo[i] * um + mdi[i] * uc + mup[i] * up1;
    b[i] = dt * s + 1.0 * mm;
  }
}
__global__ void k_set_ends(double* __restrict__ u, long long n, double bc0, double bc1) {  // time_dependent.rs:273-274
  if (threadIdx.x == 0 && blockIdx.x == 0) u[0] = bc0, u[n - 1] = bc1;
}

// ---- host orchestration --------------------------------------------------------------------------------------------
struct BigLevel {
  size_t m = 0;  // rows of the reduced system produced at this level
  double *lo = nullptr, *di = nullptr, *up = nullptr, *f = nullptr, *u = nullptr;
};
struct BigSys {
  size_t n = 0;
  std::vector<BigLevel> levels;  // levels[k]: reduced system of the k-th application of the tile kernel
  double* resid = nullptr;       // refinement residual / explicit-Euler right-hand side
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
  for (BigLevel& l : s->levels)
    for (double** p : {&l.lo, &l.di, &l.up, &l.f, &l.u})
      if (*p) cudaFree(*p), *p = nullptr;
  s->levels.clear();
  if (s->resid) cudaFree(s->resid), s->resid = nullptr;
  s->n = 0;
}

// Allocates the scratch pyramid for systems of n rows.
inline int bigsys_prepare(BigSys* s, size_t n) {
  bigsys_free(s);
  s->n = n;
  cudaError_t e;
  size_t m = n;
  while (m > (size_t)BS_MID_CAP) {
    BigLevel l;
    l.m = 2 * ((m + BS_TILE - 1) / BS_TILE);
    for (double** p : {&l.lo, &l.di, &l.up, &l.f, &l.u})
      if ((e = cudaMalloc((void**)p, l.m * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
    s->levels.push_back(l);
    m = l.m;
  }
  if ((e = cudaMalloc((void**)&s->resid, n * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
  return 0;
}

template <int P1>
inline int bigsys_mid_launch(const double* lo, const double* di, const double* up, const double* f, size_t m, double* out,
                             bool accum, cudaStream_t st) {
  constexpr size_t smem = sizeof(Pyramid<P1>);
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e1 = cudaFuncSetAttribute(k_tri_mid<P1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaError_t e2 = cudaFuncSetAttribute(k_tri_mid<P1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess) return bigsys_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute");
      attr_set = true;
    }
  }
  if (accum)
    k_tri_mid<P1, true><<<1, BS_THREADS, smem, st>>>(lo, di, up, f, (int)m, out);
  else
    k_tri_mid<P1, false><<<1, BS_THREADS, smem, st>>>(lo, di, up, f, (int)m, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : bigsys_fail(e, "k_tri_mid");
}
inline int bigsys_mid(const double* lo, const double* di, const double* up, const double* f, size_t m, double* out, bool accum,
                      cudaStream_t st, unsigned long long* launches) {
  ++*launches;
  if (m <= 32) return bigsys_mid_launch<4>(lo, di, up, f, m, out, accum, st);
  if (m <= 128) return bigsys_mid_launch<16>(lo, di, up, f, m, out, accum, st);
  if (m <= 512) return bigsys_mid_launch<64>(lo, di, up, f, m, out, accum, st);
  return bigsys_mid_launch<256>(lo, di, up, f, m, out, accum, st);
}

// out (+)= A^{-1} f for the n-row system (lo, di, up); inputs are not modified.
inline int bigsys_solve_once(BigSys* s, const double* lo, const double* di, const double* up, const double* f, double* out,
                             bool accum, cudaStream_t st, unsigned long long* launches) {
  cudaError_t e;
  if (s->levels.empty()) return bigsys_mid(lo, di, up, f, s->n, out, accum, st, launches);
  const size_t nl = s->levels.size();
  // up-sweeps
  const double *cl = lo, *cd = di, *cu = up, *cf = f;
  size_t m = s->n;
  for (size_t k = 0; k < nl; ++k) {
    BigLevel& L = s->levels[k];
    const unsigned tiles = (unsigned)((m + BS_TILE - 1) / BS_TILE);
    k_tri_tile<false, false><<<tiles, BS_THREADS, 0, st>>>(cl, cd, cu, cf, (long long)m, L.lo, L.di, L.up, L.f, nullptr, nullptr);
    ++*launches;
    if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_tri_tile<reduce>");
    cl = L.lo, cd = L.di, cu = L.up, cf = L.f, m = L.m;
  }
  {
    BigLevel& L = s->levels[nl - 1];
    if (bigsys_mid(L.lo, L.di, L.up, L.f, L.m, L.u, false, st, launches)) return 1;
  }
  // down-sweeps
  for (size_t k = nl; k-- > 0;) {
    const double *sl, *sd, *su, *sf;
    double* dst;
    size_t mm;
    bool acc = false;
    if (k == 0) sl = lo, sd = di, su = up, sf = f, dst = out, mm = s->n, acc = accum;
    else {
      BigLevel& P = s->levels[k - 1];
      sl = P.lo, sd = P.di, su = P.up, sf = P.f, dst = P.u, mm = P.m;
    }
    const unsigned tiles = (unsigned)((mm + BS_TILE - 1) / BS_TILE);
    BigLevel& L = s->levels[k];
    if (acc)
      k_tri_tile<true, true><<<tiles, BS_THREADS, 0, st>>>(sl, sd, su, sf, (long long)mm, nullptr, nullptr, nullptr, nullptr, L.u, dst);
    else
      k_tri_tile<true, false><<<tiles, BS_THREADS, 0, st>>>(sl, sd, su, sf, (long long)mm, nullptr, nullptr, nullptr, nullptr, L.u, dst);
    ++*launches;
    if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_tri_tile<solve>");
  }
  return 0;
}

inline unsigned bigsys_grid(size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = 148 * 16;
  return (unsigned)(b < cap ? b : cap);
}

// Static solve: out = K^{-1} rhs, then `refine_steps` sweeps  r = rhs - K out (double-double), out += K^{-1} r.
inline int bigsys_solve(BigSys* s, const double* lo, const double* di, const double* up, const double* rhs, double* out,
                        int refine_steps, cudaStream_t st, unsigned long long* launches) {
  if (bigsys_solve_once(s, lo, di, up, rhs, out, false, st, launches)) return 1;
  for (int k = 0; k < refine_steps; ++k) {
    k_residual_dd<<<bigsys_grid(s->n), 256, 0, st>>>(lo, di, up, rhs, out, s->resid, (long long)s->n);
    ++*launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bigsys_fail(e, "k_residual_dd");
    if (bigsys_solve_once(s, lo, di, up, s->resid, out, true, st, launches)) return 1;
  }
  return 0;
}

// One DiffussionSolverTimeDependent::solve(dt) of a single long bar (time_dependent.rs:257-280).
inline int bigsys_td_step(BigSys* s, const double* u, double* u_new, const double* mlo, const double* mdi, const double* mup,
                          const double* slo, const double* sdi, const double* sup, double dt, double bc0, double bc1,
                          cudaStream_t st, unsigned long long* launches) {
  k_td_rhs<<<bigsys_grid(s->n), 256, 0, st>>>(u, mlo, mdi, mup, slo, sdi, sup, dt, s->resid, (long long)s->n);
  ++*launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bigsys_fail(e, "k_td_rhs");
  if (bigsys_solve_once(s, mlo, mdi, mup, s->resid, u_new, false, st, launches)) return 1;
  k_set_ends<<<1, 32, 0, st>>>(u_new, (long long)s->n, bc0, bc1);
  ++*launches;
  if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_set_ends");
  return 0;
}

}  // namespace dzb
