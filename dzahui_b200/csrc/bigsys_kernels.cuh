// bigsys_kernels.cuh — one long tridiagonal system (a single bar with n >> 1024 nodes): Thomas / PCR-style hybrid.
//
// The reference solves K u = b with the serial Thomas recurrence (matrix_solver/mod.rs:17-48), which cannot use a GPU
// at n = 10^7.  Here every thread runs a *modified Thomas* sweep over R = 8 consecutive rows that eliminates the six
// interior unknowns and leaves two equations coupling the chunk's first/last unknown (u_s, u_e) with the neighbouring
// chunks — again a tridiagonal system, of a quarter of the size, in the order (s_0, e_0, s_1, e_1, ...).  The reduction
// is applied recursively:
//     level 0   registers      2048 rows per CTA (256 threads x 8 rows, 128-bit loads)
//     level 1-4 shared memory  512 -> 128 -> 32 -> 8 -> 2 rows, chunk-interleaved SoA (conflict-free)
//     grid      k_tri_tile<REDUCE> writes 2 rows per tile; k_tri_mid (one CTA, shared memory only) solves that system
//               (recursing through k_tri_tile again while it exceeds 2048 rows); k_tri_tile<SOLVE> redoes the forward
//               sweeps on chip and back-substitutes.  HBM traffic: 32 B/row (reduce) + 40 B/row (solve).
//
// Accuracy.  A plain partitioned elimination is ~150x LESS accurate than serial Thomas on the n = 10^7 stiffness matrix
// (identical chunks make identical rounding errors, which add up coherently in the nearly-cancelling row sums).  Rows are
// therefore carried in the Grassmann-Taksar-Heyman form (a, c, s, f):  a u_{i-1} + (s - a - c) u_i + c u_{i+1} = f,
// with the row-sum excess s = lo + di + up obtained once by an error-free TwoSum.  For an M-matrix (a, c <= 0 <= s) every
// update below adds quantities of one sign, so nothing cancels; for other sign patterns it is ordinary elimination.
// Measured against a __float128 Thomas of the same system (regular bar, mu = b = 1): n = 10^7 -> 1.4e-10 relative L2,
// where the reference's own f64 Thomas is at 2.6e-7 (tests/test_gpu_parity.py).
// Pivoting: none, like the reference.  The Stokes matrix (cancelling ~0 diagonal) stays on the serial kernel.
//
// One chunk, rows j = 0..7 (row 0 = s, row 7 = e), ONE forward sweep, and no division on its critical path: instead of
// the pivots beta_j the sweep carries the leading minors P_j = beta_1 ... beta_j (three-term recurrence) and rows scaled
// by P_{j-1},
//     sigma~_j = S_j P_{j-1} - A_j sigma~_{j-1},   alpha~_j = -A_j alpha~_{j-1},   delta~_j = F_j P_{j-1} - A_j delta~_{j-1},
//     g~_j = C_j P_{j-1},   P_j = sigma~_j - alpha~_j - g~_j          (row j:  alpha~_j u_s + P_j u_j + g~_j u_{j+1} = delta~_j)
// so a step is 3 dependent fp64 operations instead of ~14 (the reciprocals 1/P_j are off the chain and pipeline).  The
// chunk is first scaled by the power of two nearest 1/beta_1, which keeps P_j ~ O(1) for any magnitude of the entries.
//   e-row = (alpha~_7, g~_7, sigma~_7, delta~_7) / P_6;  the s-row needs rows 1..6 folded back into row 0, accumulated on
//   the way with G_j = (-C_0)(-C_1)...(-C_{j-1}):  s-row = (A_0, G_6 C_6 / P_6, S_0 + sum_j G_j sigma~_j / (P_j P_{j-1}),
//   F_0 + sum_j G_j delta~_j / (P_j P_{j-1})).   For an M-matrix every term of every sum has the same sign.
//   Once u_s, u_e are known: u_j = (delta~_j - alpha~_j u_s - g~_j u_{j+1}) / P_j, j = 6..1.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "tridiag_kernels.cuh"

namespace dzb {

constexpr int BS_R = 8;                       // rows per chunk at every level
constexpr int BS_THREADS = 256;               // tile kernel
constexpr int BS_TILE = BS_R * BS_THREADS;    // 2048 rows per CTA
constexpr int BS_MID_CAP = 2048;              // rows the single-CTA kernel k_tri_mid solves directly
#ifndef BS_MIN_CTAS
#define BS_MIN_CTAS 3
#endif

enum { BS_SRC_BANDS = 0, BS_SRC_ROWS = 1 };   // inputs are (lo, di, up, f) or already (a, c, s, f)
enum { BS_OPEN_LEFT = 1, BS_OPEN_RIGHT = 2 }; // the system is a block of a longer one: end couplings are kept
enum { BS_MID_SOLVE = 0, BS_MID_REDUCE = 1, BS_MID_BACKSOLVE = 2 };

// fl(lo + di + up) with a single rounding (TwoSum twice)
__device__ __forceinline__ double excess3(double lo, double di, double up) {
  const double s1 = lo + up, b1 = s1 - lo, e1 = (lo - (s1 - b1)) + (up - b1);
  const double s2 = s1 + di, b2 = s2 - s1, e2 = (s1 - (s2 - b2)) + (di - b2);
  return s2 + (e1 + e2);
}

// 2^-e for x = 1.f * 2^e (1 for zero, subnormal, inf, nan): an exact row scaling
__device__ __forceinline__ double pow2_inv(double x) {
  const int E = (__double2hiint(x) >> 20) & 0x7ff;
  return (E == 0 || E == 0x7ff) ? 1.0 : __hiloint2double((2046 - E) << 20, 0);
}

// Shared-memory level: 8*P rows, SoA, row r of the level at (r % 8) * P + r / 8.
struct LevelPtr {
  double *a, *c, *s, *d;
};
__device__ __forceinline__ int lvl_pos(int r, int P) { return (r & 7) * P + (r >> 3); }

// Ragged sizes are handled without padding rows: a level holds `rows` rows and its last chunk is simply shorter
// (len = 2..8; len == 2 has nothing to eliminate, its two rows pass through).  A single left-over row (rows % 8 == 1) is
// avoided by splitting the last nine rows 7 + 2.  So every unknown of every level is a REAL row — which is what lets
// a block of a longer bar keep its end couplings (SPIKE split across GPUs): padding rows at an open end would alias the
// neighbour's unknown instead.
__device__ __forceinline__ int chunk_len(int rows, int t) {
  const int T = (rows - 1) >> 3;  // last chunk
  if ((rows & 7) == 1 && rows > 1) return t == T ? 2 : (t == T - 1 ? 7 : (t < T ? BS_R : 0));
  return max(0, min(BS_R, rows - BS_R * t));
}
// slot (chunk-interleaved position) of row r of a level of `rows` rows with capacity P chunks
__device__ __forceinline__ int row_slot(int r, int rows, int P) {
  if ((rows & 7) == 1 && r >= rows - 2 && rows > 1) return (r - (rows - 2)) * P + ((rows - 1) >> 3);
  return lvl_pos(r, P);
}

// One chunk (thread t of P) of a shared-memory level, reduced rows -> next level (PN chunks; PN == 0: the next level is
// the final 2-row system stored at index 0/1).
template <bool FULL>  // FULL: len == 8, loops unrolled
__device__ __forceinline__ void level_reduce_impl(const LevelPtr& lv, int P, int len, const LevelPtr& nx, int PN, int t) {
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  int idx = P + t;
  double alp = lv.a[idx], sig = lv.s[idx], dlt = lv.d[idx], g = lv.c[idx];
  const double sc = (FULL || len >= 3) ? pow2_inv((sig - alp) - g) : 1.0;
  const double A0 = lv.a[t] * sc, C0 = lv.c[t] * sc, S0 = lv.s[t] * sc, F0 = lv.d[t] * sc;
  alp *= sc, sig *= sc, dlt *= sc, g *= sc;
  double s_c = C0, s_s = S0, s_f = F0;  // s-row (its a stays A0)
  double rn = 1.0;                      // normalisation of the e-row
  if (FULL || len >= 3) {
    double Pm = (sig - alp) - g;        // P_1
    double rP = __drcp_rn(Pm);
    lv.a[idx] = alp, lv.c[idx] = g, lv.s[idx] = rP, lv.d[idx] = dlt;
    double G = -C0;
    s_s = fma(G * rP, sig, S0), s_f = fma(G * rP, dlt, F0);
    double lastC = g, Cprev = g;
#pragma unroll
    for (int j = 2; j < BS_R; ++j) {
      if (FULL || j < len) {
        idx += P;
        const double Aj = lv.a[idx] * sc, Cj = lv.c[idx] * sc;
        sig = fma(lv.s[idx] * sc, Pm, -Aj * sig);
        dlt = fma(lv.d[idx] * sc, Pm, -Aj * dlt);
        alp = -Aj * alp;
        g = Cj * Pm;
        if (FULL ? j < BS_R - 1 : j < len - 1) {
          Pm = (sig - alp) - g;
          const double rPn = __drcp_rn(Pm);
          lv.a[idx] = alp, lv.c[idx] = g, lv.s[idx] = rPn, lv.d[idx] = dlt;
          G *= -Cprev;
          const double w = G * rPn * rP;
          s_s = fma(w, sig, s_s), s_f = fma(w, dlt, s_f);
          lastC = Cj, rP = rPn;
        }
        Cprev = Cj;
      }
    }
    s_c = G * lastC * rP;
    rn = rP;
  }
  nx.a[p0] = A0, nx.c[p0] = s_c, nx.s[p0] = s_s, nx.d[p0] = s_f;
  nx.a[p1] = alp * rn, nx.c[p1] = g * rn, nx.s[p1] = sig * rn, nx.d[p1] = dlt * rn;
}
__device__ __forceinline__ void level_reduce(const LevelPtr& lv, int P, int len, const LevelPtr& nx, int PN, int t) {
  if (len == BS_R) level_reduce_impl<true>(lv, P, len, nx, PN, t);
  else level_reduce_impl<false>(lv, P, len, nx, PN, t);
}
// Back-substitution of the same chunk: the next level's d[] holds the solved u_s / u_e; this level's d[] receives u.
__device__ __forceinline__ void level_backsolve(const LevelPtr& lv, int P, int len, const LevelPtr& nx, int PN, int t) {
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  const double us = nx.d[p0], ue = nx.d[p1];
  lv.d[t] = us;
  lv.d[(len - 1) * P + t] = ue;
  double nxt = ue;
  if (len == BS_R) {
#pragma unroll
    for (int j = BS_R - 2; j >= 1; --j) {
      const int idx = j * P + t;
      nxt = fma(-lv.c[idx], nxt, fma(-lv.a[idx], us, lv.d[idx])) * lv.s[idx];
      lv.d[idx] = nxt;
    }
  } else {
    for (int j = len - 2; j >= 1; --j) {
      const int idx = j * P + t;
      nxt = fma(-lv.c[idx], nxt, fma(-lv.a[idx], us, lv.d[idx])) * lv.s[idx];
      lv.d[idx] = nxt;
    }
  }
}

// Shared-memory pyramid below a level-1 system of up to 8*P1 rows: level capacities P1, ceil(P1/4), ... , 1 chunks, then the
// final 2 rows.  WARP: the pyramid belongs to one warp (tile of 256 rows) and synchronises with __syncwarp, so the warps
// of a CTA run their tiles independently; otherwise it belongs to the CTA (__syncthreads).
template <int P>
struct PyramidRows {
  static constexpr int value = 8 * P + PyramidRows<(P + 3) / 4>::value;
};
template <>
struct PyramidRows<1> {
  static constexpr int value = 8 + 2;
};
template <int P1, bool WARP = false>
struct Pyramid {
  static constexpr int ROWS = PyramidRows<P1>::value;
  double a[ROWS], c[ROWS], s[ROWS], d[ROWS];
  __device__ __forceinline__ LevelPtr level(int off) { return {a + off, c + off, s + off, d + off}; }
  __device__ __forceinline__ static void sync() {
    if (WARP) __syncwarp(); else __syncthreads();
  }
  // forward sweeps from level 1 (holding rows1 rows) to the final 2 rows; every thread of the tile must call
  __device__ __forceinline__ void reduce(int t, int rows1) {
    int off = 0, rows = rows1;
#pragma unroll
    for (int p = P1;; p = (p + 3) / 4) {
      sync();
      if (chunk_len(rows, t) > 0) level_reduce(level(off), p, chunk_len(rows, t), level(off + 8 * p), p == 1 ? 0 : (p + 3) / 4, t);
      off += 8 * p;
      rows = 2 * ((rows + BS_R - 1) / BS_R);
      if (p == 1) break;
    }
    sync();
  }
  __device__ __forceinline__ LevelPtr top() { return level(ROWS - 2); }
  // back-substitution; top().d[0..1] must hold the solved pair
  __device__ __forceinline__ void backsolve(int t, int rows1) {
    int rows_at[8], cap[8], offs[8], nl = 0;
    {
      int rows = rows1, off = 0;
#pragma unroll
      for (int p = P1;; p = (p + 3) / 4) {
        rows_at[nl] = rows, cap[nl] = p, offs[nl] = off, ++nl;
        off += 8 * p;
        rows = 2 * ((rows + BS_R - 1) / BS_R);
        if (p == 1) break;
      }
    }
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      if (k < nl) {
        const int p = cap[k];
        sync();
        if (chunk_len(rows_at[k], t) > 0)
          level_backsolve(level(offs[k]), p, chunk_len(rows_at[k], t), level(offs[k] + 8 * p), p == 1 ? 0 : (p + 3) / 4, t);
      }
    }
    sync();
  }
};

// Level 0: the forward sweep over a chunk held in registers; KEEP stores (alpha~_j, g~_j, 1/P_j, delta~_j) over
// (A, C, S, F)[1..len-2] for the back-substitution.  FULL: len == 8.
template <bool KEEP, bool FULL>
__device__ __forceinline__ void chunk0_forward(double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R], double (&F)[BS_R],
                                               int len, const LevelPtr& l1, int i0, int i1) {
  double alp = A[1], sig = S[1], dlt = F[1], g = C[1];
  const double sc = (FULL || len >= 3) ? pow2_inv((sig - alp) - g) : 1.0;
  const double A0 = A[0] * sc, C0 = C[0] * sc, S0 = S[0] * sc, F0 = F[0] * sc;
  alp *= sc, sig *= sc, dlt *= sc, g *= sc;
  double s_c = C0, s_s = S0, s_f = F0;
  double rn = 1.0;
  if (FULL || len >= 3) {
    double Pm = (sig - alp) - g;  // P_1
    double rP = __drcp_rn(Pm);
    if (KEEP) A[1] = alp, C[1] = g, S[1] = rP, F[1] = dlt;
    double G = -C0;
    s_s = fma(G * rP, sig, S0), s_f = fma(G * rP, dlt, F0);
    double lastC = g, Cprev = g;
#pragma unroll
    for (int j = 2; j < BS_R; ++j) {
      if (FULL || j < len) {
        const double Aj = A[j] * sc, Cj = C[j] * sc;
        sig = fma(S[j] * sc, Pm, -Aj * sig);
        dlt = fma(F[j] * sc, Pm, -Aj * dlt);
        alp = -Aj * alp;
        g = Cj * Pm;
        if (FULL ? j < BS_R - 1 : j < len - 1) {
          Pm = (sig - alp) - g;
          const double rPn = __drcp_rn(Pm);
          if (KEEP) A[j] = alp, C[j] = g, S[j] = rPn, F[j] = dlt;
          G *= -Cprev;
          const double w = G * rPn * rP;
          s_s = fma(w, sig, s_s), s_f = fma(w, dlt, s_f);
          lastC = Cj, rP = rPn;
        }
        Cprev = Cj;
      }
    }
    s_c = G * lastC * rP;
    rn = rP;
  }
  l1.a[i0] = A0, l1.c[i0] = s_c, l1.s[i0] = s_s, l1.d[i0] = s_f;
  l1.a[i1] = alp * rn, l1.c[i1] = g * rn, l1.s[i1] = sig * rn, l1.d[i1] = dlt * rn;
}

// ---- tile kernel --------------------------------------------------------------------------------------------------
// SOLVE == false: forward sweeps, the tile's two reduced rows -> red_{a,c,s,f}[2 tile + {0,1}].
// SOLVE == true : forward sweeps again, (u_s, u_e) of the tile from ured[2 tile + {0,1}], back-substitution, out[]
//                 (ACCUM: out[] += value, the iterative-refinement update).
// SRC: p0..p3 = (lo, di, up, f) of the caller's system, or (a, c, s, f) of a reduced system.
// TT: threads that share a tile.  256: the CTA (2048 rows, __syncthreads between the pyramid levels).  32: one warp (256
// rows, __syncwarp only), so that while one warp is inside its latency-bound pyramid the other warps of the SM keep
// streaming; a tile then leaves 2 rows per 256.
template <int SRC, bool SOLVE, bool ACCUM, int TT>
__global__ void __launch_bounds__(BS_THREADS, BS_MIN_CTAS) k_tri_tile(const double* __restrict__ p0, const double* __restrict__ p1,
                                                            const double* __restrict__ p2, const double* __restrict__ p3,
                                                            long long m, double* __restrict__ red_a,
                                                            double* __restrict__ red_c, double* __restrict__ red_s,
                                                            double* __restrict__ red_f, const double* __restrict__ ured,
                                                            double* __restrict__ out, int flags) {
  constexpr bool WARP = TT == 32;
  constexpr int TILE = BS_R * TT;
  constexpr int P1 = TT / 4;  // chunks of level 1
  __shared__ Pyramid<P1, WARP> pyrs[BS_THREADS / TT];
  Pyramid<P1, WARP>& pyr = pyrs[WARP ? threadIdx.x / 32 : 0];
  const int t = WARP ? (threadIdx.x & 31) : threadIdx.x;
  const long long tile = WARP ? (long long)blockIdx.x * (BS_THREADS / 32) + threadIdx.x / 32 : (long long)blockIdx.x;
  const long long ntiles = (m + TILE - 1) / TILE;
  if (tile >= ntiles) return;  // whole warps only (warp scope); never taken in CTA scope
  long long tile0 = tile * TILE;
  int rows0 = (int)min((long long)TILE, m - tile0);               // rows of this tile
  {  // never leave a single row to the last tile: split (TILE - 1) + 2 instead
    const long long rem = m - (ntiles - 1) * TILE;
    if (rem == 1 && ntiles > 1) {
      if (tile == ntiles - 1) tile0 -= 1, rows0 = 2;
      else if (tile == ntiles - 2) rows0 = TILE - 1;
    }
  }
  const int rows1 = 2 * ((rows0 + BS_R - 1) / BS_R);              // rows it leaves at level 1
  const int len = chunk_len(rows0, t);                            // rows of this thread's chunk (0: idle)
  const bool shifted = (rows0 & 7) == 1 && rows0 > 1 && t == ((rows0 - 1) >> 3);  // the "2" of a 7 + 2 split
  const long long g0 = tile0 + (long long)t * BS_R - (shifted ? 1 : 0);
  double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
  const bool vec = len == BS_R &&  // full chunk whose 128-bit accesses are aligned
                   !((reinterpret_cast<size_t>(p0 + g0) | reinterpret_cast<size_t>(p1 + g0) | reinterpret_cast<size_t>(p2 + g0) |
                      reinterpret_cast<size_t>(p3 + g0) | (SOLVE ? reinterpret_cast<size_t>(out + g0) : 0)) & 15);
  if (vec) {
    const double2* q0 = reinterpret_cast<const double2*>(p0 + g0);
    const double2* q1 = reinterpret_cast<const double2*>(p1 + g0);
    const double2* q2 = reinterpret_cast<const double2*>(p2 + g0);
    const double2* q3 = reinterpret_cast<const double2*>(p3 + g0);
#pragma unroll
    for (int j = 0; j < BS_R / 2; ++j) {
      const double2 v0 = __ldcs(q0 + j), v1 = __ldcs(q1 + j), v2 = __ldcs(q2 + j), v3 = __ldcs(q3 + j);
      if (SRC == BS_SRC_BANDS) {
        A[2 * j] = v0.x, A[2 * j + 1] = v0.y, S[2 * j] = v1.x, S[2 * j + 1] = v1.y;  // S holds di for now
        C[2 * j] = v2.x, C[2 * j + 1] = v2.y;
      } else {
        A[2 * j] = v0.x, A[2 * j + 1] = v0.y, C[2 * j] = v1.x, C[2 * j + 1] = v1.y;
        S[2 * j] = v2.x, S[2 * j + 1] = v2.y;
      }
      F[2 * j] = v3.x, F[2 * j + 1] = v3.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const long long g = g0 + j;
      const bool in = j < len;
      if (SRC == BS_SRC_BANDS) A[j] = in ? p0[g] : 0.0, S[j] = in ? p1[g] : 1.0, C[j] = in ? p2[g] : 0.0;
      else A[j] = in ? p0[g] : 0.0, C[j] = in ? p1[g] : 0.0, S[j] = in ? p2[g] : 1.0;
      F[j] = in ? p3[g] : 0.0;
    }
  }
  // closed ends cut the first / last coupling; a block of a longer bar (SPIKE split) keeps them
  if (g0 == 0 && len > 0 && !(flags & BS_OPEN_LEFT)) A[0] = 0.0;
  if (!(flags & BS_OPEN_RIGHT) && len > 0 && g0 + len == m) {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) C[j] = (g0 + j == m - 1) ? 0.0 : C[j];  // selects, not a dynamically indexed store
  }
  if (SRC == BS_SRC_BANDS) {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) S[j] = excess3(A[j], S[j], C[j]);
  }
  // level 0 in registers; SOLVE keeps (alpha~_j, g~_j, 1/P_j, delta~_j) in (A, C, S, F)[1..len-2]
  if (len >= 2) {
    LevelPtr l1 = pyr.level(0);
    const int i0 = lvl_pos(2 * t, P1), i1 = lvl_pos(2 * t + 1, P1);
    if (len == BS_R) chunk0_forward<SOLVE, true>(A, C, S, F, len, l1, i0, i1);
    else chunk0_forward<SOLVE, false>(A, C, S, F, len, l1, i0, i1);
  }
  pyr.reduce(t, rows1);
  LevelPtr top = pyr.top();
  if (!SOLVE) {
    if (t < 2) {
      const long long o = 2LL * tile + t;
      red_a[o] = top.a[t], red_c[o] = top.c[t], red_s[o] = top.s[t], red_f[o] = top.d[t];
    }
    return;
  }
  if (t < 2) top.d[t] = ured[2LL * tile + t];
  pyr.backsolve(t, rows1);
  if (len >= 2) {
    LevelPtr l1 = pyr.level(0);
    const double us = l1.d[lvl_pos(2 * t, P1)], ue = l1.d[lvl_pos(2 * t + 1, P1)];
    double o[BS_R];
    // A decoupled end row of the caller's system (a Dirichlet row: lo = up = 0) is returned exactly as Thomas returns it,
    // f / di (matrix_solver/mod.rs:36-39 with a zero coupling); the scaled sweeps would be off by an ulp.
    const bool fix_first = SRC == BS_SRC_BANDS && g0 == 0 && !(flags & BS_OPEN_LEFT) && p2[0] == 0.0;
    const bool fix_last = SRC == BS_SRC_BANDS && g0 + len == m && !(flags & BS_OPEN_RIGHT) && p0[m - 1] == 0.0;
    if (vec) {
      o[0] = us, o[BS_R - 1] = ue;
#pragma unroll
      for (int j = BS_R - 2; j >= 1; --j) o[j] = fma(-C[j], o[j + 1], fma(-A[j], us, F[j])) * S[j];
      if (fix_first) o[0] = p3[0] / p1[0];
      if (fix_last) o[BS_R - 1] = p3[m - 1] / p1[m - 1];
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
      double nxt = ue;
#pragma unroll
      for (int j = BS_R - 1; j >= 0; --j) {
        if (j == len - 1) o[j] = ue;
        else if (j == 0) o[j] = us;
        else if (j < len - 1) nxt = fma(-C[j], nxt, fma(-A[j], us, F[j])) * S[j], o[j] = nxt;
        else o[j] = 0.0;
      }
      if (fix_first) o[0] = p3[0] / p1[0];
#pragma unroll
      for (int j = 0; j < BS_R; ++j) {
        if (fix_last && j == len - 1) o[j] = p3[m - 1] / p1[m - 1];
        if (j < len) out[g0 + j] = ACCUM ? out[g0 + j] + o[j] : o[j];
      }
    }
  }
}

// ---- single-CTA kernel: solves an m-row system (m <= 8 P1 <= 2048) entirely in shared memory ------------------------
// The rows are loaded coalesced straight into level 1 of the pyramid; inputs are not modified.  P1 in {4, 16, 64, 256}: 32 / 128 / 512 / 2048 rows.
// MODE: BS_MID_SOLVE      closed system -> out[0..m)
//       BS_MID_REDUCE     open block: the two rows left after eliminating everything inside -> top8 = (a,c,s,f) x 2
//       BS_MID_BACKSOLVE  open block whose end unknowns (us, ue) are now known -> out[0..m)
template <int P1, int SRC, int MODE, bool ACCUM>
__global__ void __launch_bounds__(BS_THREADS) k_tri_mid(const double* __restrict__ p0, const double* __restrict__ p1,
                                                        const double* __restrict__ p2, const double* __restrict__ p3, int m,
                                                        double* __restrict__ out, int flags, double* __restrict__ top8,
                                                        double us, double ue) {
  extern __shared__ double4 mid_smem[];
  Pyramid<P1>& pyr = *reinterpret_cast<Pyramid<P1>*>(mid_smem);
  const int t = threadIdx.x;
  const bool open_left = flags & BS_OPEN_LEFT, open_right = flags & BS_OPEN_RIGHT;
  LevelPtr l1 = pyr.level(0);
  for (int r = t; r < m; r += BS_THREADS) {
    const int p = row_slot(r, m, P1);
    const double a = (r > 0 || open_left) ? p0[r] : 0.0;
    const double c = (r < m - 1 || open_right) ? (SRC == BS_SRC_BANDS ? p2[r] : p1[r]) : 0.0;
    l1.a[p] = a, l1.c[p] = c, l1.s[p] = SRC == BS_SRC_BANDS ? excess3(a, p1[r], c) : p2[r], l1.d[p] = p3[r];
  }
  pyr.reduce(t, m);
  LevelPtr top = pyr.top();
  if (MODE == BS_MID_REDUCE) {
    if (t < 2) top8[4 * t + 0] = top.a[t], top8[4 * t + 1] = top.c[t], top8[4 * t + 2] = top.s[t], top8[4 * t + 3] = top.d[t];
    return;
  }
  if (t == 0) {
    if (MODE == BS_MID_SOLVE) {  // closed 2 x 2 system (a_0 = c_1 = 0): [s0 - c0, c0; a1, s1 - a1] (u_s, u_e) = (f0, f1)
      const double c0 = top.c[0], s0 = top.s[0], f0 = top.d[0], a1 = top.a[1], s1 = top.s[1], f1 = top.d[1];
      const double det = (s0 * s1 - s0 * a1) - c0 * s1;  // = b0 b1 - c0 a1 without the cancelling c0 a1 terms
      top.d[0] = ((s1 - a1) * f0 - c0 * f1) / det;
      top.d[1] = ((s0 - c0) * f1 - a1 * f0) / det;
    } else if (top8) {  // the pair the interface solve left in device memory
      top.d[0] = top8[0], top.d[1] = top8[1];
    } else {
      top.d[0] = us, top.d[1] = ue;
    }
  }
  pyr.backsolve(t, m);
  for (int r = t; r < m; r += BS_THREADS) {
    double v = l1.d[row_slot(r, m, P1)];
    if (SRC == BS_SRC_BANDS) {  // exact Dirichlet end rows, as in k_tri_tile
      if (r == 0 && !open_left && p2[0] == 0.0) v = p3[0] / p1[0];
      if (r == m - 1 && !open_right && p0[m - 1] == 0.0) v = p3[m - 1] / p1[m - 1];
    }
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
                                                     const double* __restrict__ u, double* __restrict__ r, long long n,
                                                     int flags, double u_left, double u_right,
                                                     const double* __restrict__ d_left,
                                                     const double* __restrict__ d_right) {
  if (d_left) u_left = *d_left;     // halo values left in device memory by the edge exchange
  if (d_right) u_right = *d_right;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = f[i], comp = 0.0;
    if (i > 0) dd_sub_prod(s, comp, lo[i], u[i - 1]);
    else if (flags & BS_OPEN_LEFT) dd_sub_prod(s, comp, lo[i], u_left);     // halo value of the previous block
    dd_sub_prod(s, comp, di[i], u[i]);
    if (i + 1 < n) dd_sub_prod(s, comp, up[i], u[i + 1]);
    else if (flags & BS_OPEN_RIGHT) dd_sub_prod(s, comp, up[i], u_right);   // halo value of the next block
    r[i] = s + comp;
  }
}

// ---- SPIKE interface system on the device: rows[8 r ..] = block r's (a, c, s, f) x 2, 2 n_blocks unknowns ----------
// Serial Thomas in row-excess form (sigma'_i = s_i + m sigma'_{i-1}, beta_i = sigma'_i - c_i: same-signed sums for an
// M-matrix); every rank runs it redundantly on the gathered rows, so nothing but the all-gather crosses NVLink.
__global__ void k_interface_solve(const double* __restrict__ rows, int m, double* __restrict__ ends, double* __restrict__ work) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sig = rows[2] - rows[0];  // row 0: the (zero) outer coupling of the closed end leaves the excess
  double beta = sig - rows[1], fp = rows[3];
  work[0] = beta, work[m] = fp;
  for (int i = 1; i < m; ++i) {
    const double a = rows[4 * i], c = i == m - 1 ? 0.0 : rows[4 * i + 1], s = rows[4 * i + 2], f = rows[4 * i + 3];
    const double mm = -a / beta;
    const double s_eff = i == m - 1 ? s - rows[4 * i + 1] : s;  // drop the stored (zero) outer coupling from the excess
    sig = fma(mm, sig, s_eff);
    fp = fma(mm, fp, f);
    beta = sig - c;
    work[i] = beta, work[m + i] = fp;
  }
  double x = work[2 * m - 1] / work[m - 1];
  ends[m - 1] = x;
  for (int i = m - 2; i >= 0; --i) {
    x = (work[m + i] - rows[4 * i + 1] * x) / work[i];
    ends[i] = x;
  }
}

// ---- is the system one the partitioned elimination is reliable on? -------------------------------------------------
// The chunk sweeps eliminate every chunk on its own (local Dirichlet problems).  That is as stable as serial Thomas
// for an M-matrix (lo, up <= 0 and the row sum not below -1e-6 of the diagonal: weak dominance up to assembly noise) and
// for a strictly dominant matrix with positive couplings (the mass matrix); on an indefinite system (e.g. the reference's
// kink-straddling quadrature on a strongly irregular mesh, or a cell Peclet number above 1) the local problems can be
// nearly singular while the serial recurrence, which always carries the left boundary along, is not.
// flags: bit 0 set = not M-like, bit 1 set = not mass-like.  Interior rows only.
__global__ void __launch_bounds__(256) k_tri_classify(const double* __restrict__ lo, const double* __restrict__ di,
                                                      const double* __restrict__ up, long long n, int* __restrict__ flags) {
  int f = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n - 1; i += (long long)gridDim.x * blockDim.x) {
    const double a = lo[i], b = di[i], c = up[i];
    if (!(a <= 0.0 && c <= 0.0 && excess3(a, b, c) >= -1e-6 * b)) f |= 1;
    if (!(a >= 0.0 && c >= 0.0 && b >= 1.25 * (a + c))) f |= 2;
  }
  if (f) atomicOr(flags, f);
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
  double *a = nullptr, *c = nullptr, *s = nullptr, *f = nullptr, *u = nullptr;
};
struct BigSys {
  size_t n = 0;
  int tt = 256;  // threads per tile of k_tri_tile: 256 (CTA tiles of 2048 rows; measured 4 % faster) or 32 (warp tiles of 256 rows)
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
    for (double** p : {&l.a, &l.c, &l.s, &l.f, &l.u})
      if (*p) cudaFree(*p), *p = nullptr;
  s->levels.clear();
  if (s->resid) cudaFree(s->resid), s->resid = nullptr;
  s->n = 0;
}

// Allocates the scratch pyramid for systems of n rows.
inline int bigsys_prepare(BigSys* s, size_t n) {
  bigsys_free(s);
  s->n = n;
  if (const char* v = std::getenv("DZB_TILE_THREADS")) s->tt = std::atoi(v) == 32 ? 32 : 256;  // tuning knob
  cudaError_t e;
  size_t m = n;
  while (m > (size_t)BS_MID_CAP) {
    BigLevel l;
    const size_t tile = (size_t)BS_R * (size_t)s->tt;
    l.m = 2 * ((m + tile - 1) / tile);
    for (double** p : {&l.a, &l.c, &l.s, &l.f, &l.u})
      if ((e = cudaMalloc((void**)p, l.m * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
    s->levels.push_back(l);
    m = l.m;
  }
  if ((e = cudaMalloc((void**)&s->resid, n * sizeof(double))) != cudaSuccess) return bigsys_fail(e, "cudaMalloc");
  return 0;
}

template <int P1, int SRC, int MODE>
inline int bigsys_mid_launch(const double* p0, const double* p1, const double* p2, const double* p3, size_t m, double* out,
                             bool accum, int flags, double* top8, double us, double ue, cudaStream_t st) {
  constexpr size_t smem = sizeof(Pyramid<P1>);
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e1 = cudaFuncSetAttribute(k_tri_mid<P1, SRC, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaError_t e2 = cudaFuncSetAttribute(k_tri_mid<P1, SRC, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess) return bigsys_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute");
      attr_set = true;
    }
  }
  if (accum)
    k_tri_mid<P1, SRC, MODE, true><<<1, BS_THREADS, smem, st>>>(p0, p1, p2, p3, (int)m, out, flags, top8, us, ue);
  else
    k_tri_mid<P1, SRC, MODE, false><<<1, BS_THREADS, smem, st>>>(p0, p1, p2, p3, (int)m, out, flags, top8, us, ue);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : bigsys_fail(e, "k_tri_mid");
}
template <int SRC, int MODE>
inline int bigsys_mid(const double* p0, const double* p1, const double* p2, const double* p3, size_t m, double* out, bool accum,
                      int flags, double* top8, double us, double ue, cudaStream_t st, unsigned long long* launches) {
  ++*launches;
  if (m <= 32) return bigsys_mid_launch<4, SRC, MODE>(p0, p1, p2, p3, m, out, accum, flags, top8, us, ue, st);
  if (m <= 128) return bigsys_mid_launch<16, SRC, MODE>(p0, p1, p2, p3, m, out, accum, flags, top8, us, ue, st);
  if (m <= 512) return bigsys_mid_launch<64, SRC, MODE>(p0, p1, p2, p3, m, out, accum, flags, top8, us, ue, st);
  return bigsys_mid_launch<256, SRC, MODE>(p0, p1, p2, p3, m, out, accum, flags, top8, us, ue, st);
}

template <int SRC, bool SOLVE, bool ACCUM>
inline void bigsys_launch_tile(int tt, size_t m, const double* p0, const double* p1, const double* p2, const double* p3, double* ra,
                               double* rc, double* rs, double* rf, const double* ured, double* out, int flags, cudaStream_t st) {
  if (tt == 32) {
    const size_t tiles = (m + 255) / 256;
    k_tri_tile<SRC, SOLVE, ACCUM, 32><<<(unsigned)((tiles + 7) / 8), BS_THREADS, 0, st>>>(p0, p1, p2, p3, (long long)m, ra, rc, rs, rf, ured, out, flags);
  } else {
    const size_t tiles = (m + BS_TILE - 1) / BS_TILE;
    k_tri_tile<SRC, SOLVE, ACCUM, 256><<<(unsigned)tiles, BS_THREADS, 0, st>>>(p0, p1, p2, p3, (long long)m, ra, rc, rs, rf, ured, out, flags);
  }
}

// forward sweeps of every grid level: each application of k_tri_tile leaves 2 rows per tile
inline int bigsys_forward(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                          cudaStream_t st, unsigned long long* launches) {
  cudaError_t e;
  for (size_t k = 0; k < s->levels.size(); ++k) {
    BigLevel& L = s->levels[k];
    if (k == 0) {
      bigsys_launch_tile<BS_SRC_BANDS, false, false>(s->tt, s->n, lo, di, up, f, L.a, L.c, L.s, L.f, nullptr, nullptr, flags, st);
    } else {
      BigLevel& Q = s->levels[k - 1];
      bigsys_launch_tile<BS_SRC_ROWS, false, false>(s->tt, Q.m, Q.a, Q.c, Q.s, Q.f, L.a, L.c, L.s, L.f, nullptr, nullptr, flags, st);
    }
    ++*launches;
    if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_tri_tile<reduce>");
  }
  return 0;
}
// back-substitution of every grid level; levels.back().u must hold the solved top-level unknowns
inline int bigsys_backward(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                           double* out, bool accum, cudaStream_t st, unsigned long long* launches) {
  cudaError_t e;
  for (size_t k = s->levels.size(); k-- > 0;) {
    BigLevel& L = s->levels[k];
    if (k == 0) {
      if (accum)
        bigsys_launch_tile<BS_SRC_BANDS, true, true>(s->tt, s->n, lo, di, up, f, nullptr, nullptr, nullptr, nullptr, L.u, out, flags, st);
      else
        bigsys_launch_tile<BS_SRC_BANDS, true, false>(s->tt, s->n, lo, di, up, f, nullptr, nullptr, nullptr, nullptr, L.u, out, flags, st);
    } else {
      BigLevel& Q = s->levels[k - 1];
      bigsys_launch_tile<BS_SRC_ROWS, true, false>(s->tt, Q.m, Q.a, Q.c, Q.s, Q.f, nullptr, nullptr, nullptr, nullptr, L.u, Q.u, flags, st);
    }
    ++*launches;
    if ((e = cudaGetLastError()) != cudaSuccess) return bigsys_fail(e, "k_tri_tile<solve>");
  }
  return 0;
}

// out (+)= A^{-1} f for the closed n-row system (lo, di, up); inputs are not modified.
inline int bigsys_solve_once(BigSys* s, const double* lo, const double* di, const double* up, const double* f, double* out,
                             bool accum, cudaStream_t st, unsigned long long* launches) {
  if (s->levels.empty())
    return bigsys_mid<BS_SRC_BANDS, BS_MID_SOLVE>(lo, di, up, f, s->n, out, accum, 0, nullptr, 0.0, 0.0, st, launches);
  if (bigsys_forward(s, lo, di, up, f, 0, st, launches)) return 1;
  BigLevel& L = s->levels.back();
  if (bigsys_mid<BS_SRC_ROWS, BS_MID_SOLVE>(L.a, L.c, L.s, L.f, L.m, L.u, false, 0, nullptr, 0.0, 0.0, st, launches)) return 1;
  return bigsys_backward(s, lo, di, up, f, 0, out, accum, st, launches);
}

// SPIKE split, first half: eliminate everything inside this block; top8 (device, 8 doubles) receives its two interface
// rows (a, c, s, f) x 2: a couples the block's first unknown to the previous block's last, c the last to the next's first.
inline int bigsys_block_reduce(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                               double* top8, cudaStream_t st, unsigned long long* launches) {
  if (s->levels.empty())
    return bigsys_mid<BS_SRC_BANDS, BS_MID_REDUCE>(lo, di, up, f, s->n, nullptr, false, flags, top8, 0.0, 0.0, st, launches);
  if (bigsys_forward(s, lo, di, up, f, flags, st, launches)) return 1;
  BigLevel& L = s->levels.back();
  return bigsys_mid<BS_SRC_ROWS, BS_MID_REDUCE>(L.a, L.c, L.s, L.f, L.m, nullptr, false, flags, top8, 0.0, 0.0, st, launches);
}
// SPIKE split, second half: the interface solve gave this block's end unknowns (us, ue); back-substitute.
inline int bigsys_block_backsolve(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                                  double us, double ue, const double* d_ends, double* out, bool accum, cudaStream_t st,
                                  unsigned long long* launches) {
  double* de = const_cast<double*>(d_ends);  // device pair (us, ue); takes precedence over the host values when non-null
  if (s->levels.empty())
    return bigsys_mid<BS_SRC_BANDS, BS_MID_BACKSOLVE>(lo, di, up, f, s->n, out, accum, flags, de, us, ue, st, launches);
  BigLevel& L = s->levels.back();
  if (bigsys_mid<BS_SRC_ROWS, BS_MID_BACKSOLVE>(L.a, L.c, L.s, L.f, L.m, L.u, false, flags, de, us, ue, st, launches)) return 1;
  return bigsys_backward(s, lo, di, up, f, flags, out, accum, st, launches);
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
    k_residual_dd<<<bigsys_grid(s->n), 256, 0, st>>>(lo, di, up, rhs, out, s->resid, (long long)s->n, 0, 0.0, 0.0, nullptr, nullptr);
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
