// bigsys_kernels.cuh — one long tridiagonal system (a single bar with n >> 1024 nodes): Thomas / PCR-style hybrid.
//
// The reference solves K u = b with the serial Thomas recurrence (matrix_solver/mod.rs:17-48), which cannot use a GPU
// at n = 10^7.  Here every thread runs a *modified Thomas* sweep over R = 8 consecutive rows that eliminates the six
// interior unknowns and leaves two equations coupling the chunk's first/last unknown (u_s, u_e) with the neighbouring
// chunks — again a tridiagonal system, of a quarter of the size, in the order (s_0, e_0, s_1, e_1, ...).  The reduction
// is applied recursively:
//     level 0   registers      2048 rows per CTA (256 threads x 8 rows; TMA-staged through shared memory, or direct
//                              128-bit loads for ragged / unaligned tiles)
//     level 1-4 shared memory  512 -> 128 -> 32 -> 8 -> 2 rows, chunk-interleaved SoA with a padded row stride
//     grid      k_tri_tile[_tma]<REDUCE> hands 2 rows per tile (full pyramid) or, when the application is throughput
//               bound, the 128 rows of level 2 (one shared-memory level only) to the next application; k_tri_mid (one
//               CTA, shared memory only) solves what is left once it is <= 2048 rows; k_tri_tile[_tma]<SOLVE> redoes
//               the forward sweeps on chip and back-substitutes.  HBM traffic: 32 B/row (reduce) + 40 B/row (solve).
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
// The minors themselves are advanced by the equivalent two-term form  P_j = (S_j - C_j) P_{j-1} - A_j D_{j-1},
// D_j = sigma~_j - alpha~_j = S_j P_{j-1} - A_j D_{j-1}  (same-signed sums again), so a step has 2 dependent fp64 operations on
// its chain instead of ~14; the reciprocals 1/P_j (hardware seed + two Newton steps, pivot_rcp) are off the chain.  The
// chunk is first scaled by the power of two nearest 1/beta_1, which keeps P_j ~ O(1) for any magnitude of the entries.
//   e-row = (alpha~_7, g~_7, sigma~_7, delta~_7) / P_6;  the s-row needs rows 1..6 folded back into row 0, accumulated on
//   the way with G_j = (-C_0)(-C_1)...(-C_{j-1}):  s-row = (A_0, G_6 C_6 / P_6, S_0 + sum_j G_j sigma~_j / (P_j P_{j-1}),
//   F_0 + sum_j G_j delta~_j / (P_j P_{j-1})).   For an M-matrix every term of every sum has the same sign.
//   Once u_s, u_e are known: u_j = (delta~_j - alpha~_j u_s - g~_j u_{j+1}) / P_j, j = 6..1.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "comm_kernels.cuh"
#include "tridiag_kernels.cuh"

namespace dzb {

constexpr int BS_R = 8;                       // rows per chunk at every level
constexpr int BS_THREADS = 256;               // tile kernel
constexpr int BS_TILE = BS_R * BS_THREADS;    // 2048 rows per CTA
constexpr int BS_MID_CAP = 2048;              // rows the single-CTA kernel k_tri_mid solves directly
constexpr int BS_TMA_MIN_TILES = 64;          // smaller systems are launch-latency bound: plain k_tri_tile
constexpr int BS_FUSED_MAX_TILES = 512;       // most tiles k_tri_fused's in-kernel middle solve has room for
constexpr int BS_SHALLOW_MIN_TILES = 445;     // more tiles than one cooperative launch can hold (3 per SM): throughput bound from here on
#ifndef BS_MIN_CTAS
#define BS_MIN_CTAS 3
#endif

enum { BS_SRC_BANDS = 0, BS_SRC_ROWS = 1, BS_SRC_NODES = 2 };  // inputs are (lo, di, up, f), already (a, c, s, f), or rows formed on chip from node coordinates (onelaunch_kernels.cuh)
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

// 1/x for the pivots of the sweeps: the hardware seed (MUFU.RCP64H, ~2^-23) and two Newton steps, no special cases and
// therefore no branch — __drcp_rn's slow-path check is a scheduling barrier that serialised the unrolled row-steps (each
// 8-row sweep took ~1500 cycles).  Within an ulp of the correctly rounded value; the pivots are O(1) by construction
// (chunks are pre-scaled by a power of two), so the flushed denormal range is never reached on a system k_tri_classify
// accepts, and a zero pivot gives NaN instead of inf — garbage either way.
__device__ __forceinline__ double pivot_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Shared-memory level of capacity p chunks: SoA, row r at (r % 8) * P + r / 8 with the row stride P = lvl_stride(p) = p + 2.
// Chunk t reads its rows at j P + t (consecutive threads, consecutive words); the two pad words make the OTHER access
// pattern conflict-free as well — thread t writing rows 2t, 2t+1 of the next level, or reading row t for the copy-out —
// since (r % 8) * (p + 2) spreads the four even (odd) rows a half-warp touches over all 16 64-bit banks.  With P = p it
// was a 4-way conflict on every level-1 store (ncu: 4.0 wavefronts per ideal one).
struct LevelPtr {
  double *a, *c, *s, *d;
};
__host__ __device__ constexpr int lvl_stride(int p) { return p + 2; }
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
// slot (chunk-interleaved position) of row r of a level of `rows` rows with row stride P
__device__ __forceinline__ int row_slot(int r, int rows, int P) {
  if ((rows & 7) == 1 && r >= rows - 2 && rows > 1) return (r - (rows - 2)) * P + ((rows - 1) >> 3);
  return lvl_pos(r, P);
}

// One chunk (thread t) of a shared-memory level with row stride P, reduced rows -> next level (row stride PN; PN == 0: the
// next level is the final 2-row system stored at index 0/1).
// PRE: the chunk's rows are all read before the sweep starts.  With KEEP every row-step ends in stores into the very arrays
// the next step reads, and when the row stride is a runtime value the compiler cannot tell them apart: each step's loads then
// wait for the previous step's stores, reciprocal included (~125 instead of ~25 cycles per step).  Costs 64 registers.
template <bool FULL, bool KEEP, bool PRE = false>  // FULL: len == 8, loops unrolled.  KEEP: leave (alpha~, g~, 1/P, delta~) for level_backsolve
__device__ __forceinline__ void level_reduce_impl(const LevelPtr& lv, int P, int len, const LevelPtr& nx, int PN, int t) {
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  double ra[BS_R], rc[BS_R], rs[BS_R], rd[BS_R];
  if (PRE) {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const bool in = FULL || j < len;
      const int q = j * P + t;
      ra[j] = in ? lv.a[q] : 0.0, rc[j] = in ? lv.c[q] : 0.0, rs[j] = in ? lv.s[q] : 1.0, rd[j] = in ? lv.d[q] : 0.0;
    }
  }
  int idx = P + t;
  double alp = PRE ? ra[1] : lv.a[idx], sig = PRE ? rs[1] : lv.s[idx], dlt = PRE ? rd[1] : lv.d[idx], g = PRE ? rc[1] : lv.c[idx];
  const double sc = (FULL || len >= 3) ? pow2_inv((sig - alp) - g) : 1.0;
  const double A0 = (PRE ? ra[0] : lv.a[t]) * sc, C0 = (PRE ? rc[0] : lv.c[t]) * sc, S0 = (PRE ? rs[0] : lv.s[t]) * sc,
               F0 = (PRE ? rd[0] : lv.d[t]) * sc;
  alp *= sc, sig *= sc, dlt *= sc, g *= sc;
  double s_c = C0, s_s = S0, s_f = F0;  // s-row (its a stays A0)
  double rn = 1.0;                      // normalisation of the e-row
  if (FULL || len >= 3) {
    double D = sig - alp;               // sigma~ - alpha~
    double Pm = D - g;                  // P_1
    double rP = pivot_rcp(Pm);
    if (KEEP) lv.a[idx] = alp, lv.c[idx] = g, lv.s[idx] = rP, lv.d[idx] = dlt;
    double G = -C0;
    s_s = fma(G * rP, sig, S0), s_f = fma(G * rP, dlt, F0);
    double lastC = g, Cprev = g;
#pragma unroll
    for (int j = 2; j < BS_R; ++j) {
      if (FULL || j < len) {
        idx += P;
        // the chunk's power-of-two scale rides on P_{j-1} (exact), so only A_j needs its own scaling
        const double Ps = Pm * sc, nA = -((PRE ? ra[j] : lv.a[idx]) * sc), Cr = PRE ? rc[j] : lv.c[idx], Sr = PRE ? rs[j] : lv.s[idx];
        const double q = nA * D;
        const double Pn = fma(Sr - Cr, Ps, q);  // P_j by its own two-term recurrence (see chunk0_forward)
        D = fma(Sr, Ps, q);
        sig = fma(Sr, Ps, nA * sig);
        dlt = fma(PRE ? rd[j] : lv.d[idx], Ps, nA * dlt);
        alp = nA * alp;
        g = Cr * Ps;
        if (FULL ? j < BS_R - 1 : j < len - 1) {
          Pm = Pn;
          const double rPn = pivot_rcp(Pm);
          if (KEEP) lv.a[idx] = alp, lv.c[idx] = g, lv.s[idx] = rPn, lv.d[idx] = dlt;
          G *= -Cprev;
          const double w = G * rPn * rP;
          s_s = fma(w, sig, s_s), s_f = fma(w, dlt, s_f);
          lastC = Cr * sc, rP = rPn;
        }
        Cprev = Cr * sc;
      }
    }
    s_c = G * lastC * rP;
    rn = rP;
  }
  nx.a[p0] = A0, nx.c[p0] = s_c, nx.s[p0] = s_s, nx.d[p0] = s_f;
  nx.a[p1] = alp * rn, nx.c[p1] = g * rn, nx.s[p1] = sig * rn, nx.d[p1] = dlt * rn;
}
template <bool KEEP>
__device__ __forceinline__ void level_reduce(const LevelPtr& lv, int P, int len, const LevelPtr& nx, int PN, int t) {
  if (len == BS_R) level_reduce_impl<true, KEEP>(lv, P, len, nx, PN, t);
  else level_reduce_impl<false, KEEP>(lv, P, len, nx, PN, t);
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

// Out-of-line copies of the two level routines for pyramids that run once per launch (the per-CTA and the middle system of
// the one-launch kernel): the unrolled per-level instantiations reduce_to / backsolve_from make of them are ~7000 instructions
// of code executed once, and fetching that cost more than executing it (measured: 19 us for the 592-row middle solve).
__device__ __noinline__ void level_reduce_cold(LevelPtr lv, int P, int len, LevelPtr nx, int PN, int t) {
  level_reduce_impl<false, true, true>(lv, P, len, nx, PN, t);
}
__device__ __noinline__ void level_backsolve_cold(LevelPtr lv, int P, int len, LevelPtr nx, int PN, int t) {
  const int p0 = PN ? lvl_pos(2 * t, PN) : 0, p1 = PN ? lvl_pos(2 * t + 1, PN) : 1;
  const double us = nx.d[p0], ue = nx.d[p1];
  double a[BS_R], c[BS_R], d[BS_R], r[BS_R];  // every multiplier before the first store (see level_reduce_impl, PRE)
#pragma unroll
  for (int j = 1; j < BS_R - 1; ++j) {
    const bool in = j < len - 1;
    const int q = j * P + t;
    a[j] = in ? lv.a[q] : 0.0, c[j] = in ? lv.c[q] : 0.0, d[j] = in ? lv.d[q] : 0.0, r[j] = in ? lv.s[q] : 0.0;
  }
  lv.d[t] = us;
  lv.d[(len - 1) * P + t] = ue;
  double nxt = ue;
#pragma unroll
  for (int j = BS_R - 2; j >= 1; --j) {
    if (j < len - 1) {
      nxt = fma(-c[j], nxt, fma(-a[j], us, d[j])) * r[j];
      lv.d[j * P + t] = nxt;
    }
  }
}

// Shared-memory pyramid below a level-1 system of up to 8*P1 rows: level capacities P1, ceil(P1/4), ... , 1 chunks, then the
// final 2 rows.  It belongs to the CTA: the levels are separated by __syncthreads.
template <int P>
struct PyramidRows {
  static constexpr int value = 8 * lvl_stride(P) + PyramidRows<(P + 3) / 4>::value;
};
template <>
struct PyramidRows<1> {
  static constexpr int value = 8 * lvl_stride(1) + 2;
};
template <int P1>
struct Pyramid {
  static constexpr int ROWS = PyramidRows<P1>::value;
  double a[ROWS], c[ROWS], s[ROWS], d[ROWS];
  __device__ __forceinline__ LevelPtr level(int off) { return {a + off, c + off, s + off, d + off}; }
  __device__ __forceinline__ static void sync() { __syncthreads(); }
  // Forward sweeps from level 1 (holding rows1 rows); every thread of the tile must call.  NLEV == 0: down to the final 2
  // rows.  NLEV > 0: only NLEV level reductions.  The level that is left starts at row `off` of the arrays, holds `rows`
  // rows and is chunk-interleaved with row stride `cap` (row r at off + lvl_pos(r, cap); cap == 0: the final pair, at off + r).
  struct Left {
    int off, rows, cap;
    __device__ __forceinline__ int pos(int r) const { return off + (cap ? lvl_pos(r, cap) : r); }
  };
  template <int NLEV, bool KEEP = true>  // KEEP == false: no back-substitution will follow
  __device__ __forceinline__ Left reduce_to(int t, int rows1) {
    Left out{0, rows1, lvl_stride(P1)};
    int lev = 0;
#pragma unroll
    for (int p = P1;; p = (p + 3) / 4) {
      sync();
      const int pn = p == 1 ? 0 : lvl_stride((p + 3) / 4);  // row stride of the next level
      if (chunk_len(out.rows, t) > 0)
        level_reduce<KEEP>(level(out.off), lvl_stride(p), chunk_len(out.rows, t), level(out.off + 8 * lvl_stride(p)), pn, t);
      out.off += 8 * lvl_stride(p);
      out.rows = 2 * ((out.rows + BS_R - 1) / BS_R);
      out.cap = pn;
      ++lev;
      if (p == 1 || lev == NLEV) break;
    }
    sync();
    return out;
  }
  __device__ __forceinline__ void reduce(int t, int rows1) { reduce_to<0>(t, rows1); }
  __device__ __forceinline__ LevelPtr top() { return level(ROWS - 2); }
  // back-substitution through the levels reduce_to<NLEV> went through; the d[] of the level it left must hold that
  // level's solved unknowns
  template <int NLEV>
  __device__ __forceinline__ void backsolve_from(int t, int rows1) {
    int rows_at[8], cap[8], offs[8], nl = 0;
    {
      int rows = rows1, off = 0;
#pragma unroll
      for (int p = P1;; p = (p + 3) / 4) {
        rows_at[nl] = rows, cap[nl] = p, offs[nl] = off, ++nl;
        off += 8 * lvl_stride(p);
        rows = 2 * ((rows + BS_R - 1) / BS_R);
        if (p == 1 || nl == NLEV) break;
      }
    }
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      if (k < nl) {
        const int p = cap[k];
        sync();
        if (chunk_len(rows_at[k], t) > 0)
          level_backsolve(level(offs[k]), lvl_stride(p), chunk_len(rows_at[k], t), level(offs[k] + 8 * lvl_stride(p)),
                          p == 1 ? 0 : lvl_stride((p + 3) / 4), t);
      }
    }
    sync();
  }
  __device__ __forceinline__ void backsolve(int t, int rows1) { backsolve_from<0>(t, rows1); }
  // reduce / backsolve with ONE copy of the level code and a runtime loop over the levels (see level_reduce_cold)
  __device__ __forceinline__ void reduce_cold(int t, int rows1) {
    int off = 0, rows = rows1;
#pragma unroll 1
    for (int p = P1;; p = (p + 3) / 4) {
      sync();
      const int len = chunk_len(rows, t);
      if (len > 0) level_reduce_cold(level(off), lvl_stride(p), len, level(off + 8 * lvl_stride(p)), p == 1 ? 0 : lvl_stride((p + 3) / 4), t);
      off += 8 * lvl_stride(p);
      rows = 2 * ((rows + BS_R - 1) / BS_R);
      if (p == 1) break;
    }
    sync();
  }
  __device__ __forceinline__ void backsolve_cold(int t, int rows1) {
    int nl = 1;
    for (int p = P1; p != 1; p = (p + 3) / 4) ++nl;
#pragma unroll 1
    for (int k = nl - 1; k >= 0; --k) {
      int p = P1, off = 0, rows = rows1;  // geometry of level k
      for (int q = 0; q < k; ++q) off += 8 * lvl_stride(p), rows = 2 * ((rows + BS_R - 1) / BS_R), p = (p + 3) / 4;
      sync();
      const int len = chunk_len(rows, t);
      if (len > 0) level_backsolve_cold(level(off), lvl_stride(p), len, level(off + 8 * lvl_stride(p)), p == 1 ? 0 : lvl_stride((p + 3) / 4), t);
    }
    sync();
  }
};
// rows a tile holding rows1 level-1 rows leaves after NLEV level reductions (NLEV == 0: all of them)
__host__ __device__ constexpr int bs_rows_left_of(int rows1, int P1, int NLEV) {
  int rows = rows1;
  for (int p = P1, lev = 0;; p = (p + 3) / 4) {
    rows = 2 * ((rows + BS_R - 1) / BS_R);
    ++lev;
    if (p == 1 || lev == NLEV) break;
  }
  return rows;
}
// rows a full tile of 8 P1 level-1 rows leaves after NLEV level reductions (NLEV == 0: all of them)
__host__ __device__ constexpr int bs_rows_left(int P1, int NLEV) {
  int rows = 8 * P1;
  for (int p = P1, lev = 0;; p = (p + 3) / 4) {
    rows = 2 * ((rows + BS_R - 1) / BS_R);
    ++lev;
    if (p == 1 || lev == NLEV) break;
  }
  return rows;
}

// Level 0: the forward sweep over a chunk held in registers; KEEP stores (alpha~_j, g~_j, 1/P_j, delta~_j) over
// (A, C, S, F)[1..len-2] for the back-substitution.  FULL: len == 8.
// chunk0_sweep leaves the chunk's two reduced rows in o = (a, c, s, f) of the s-row, then of the e-row.
template <bool KEEP, bool FULL>
__device__ __forceinline__ void chunk0_sweep(double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R], double (&F)[BS_R], int len,
                                             double (&o)[8]) {
  double alp = A[1], sig = S[1], dlt = F[1], g = C[1];
  const double sc = (FULL || len >= 3) ? pow2_inv((sig - alp) - g) : 1.0;
  const double A0 = A[0] * sc, C0 = C[0] * sc, S0 = S[0] * sc, F0 = F[0] * sc;
  alp *= sc, sig *= sc, dlt *= sc, g *= sc;
  double s_c = C0, s_s = S0, s_f = F0;
  double rn = 1.0;
  if (FULL || len >= 3) {
    // The minors are not formed as sigma~_j - alpha~_j - g~_j (four dependent operations per row-step) but by the
    // equivalent two-term recurrence  P_j = (S_j - C_j) P_{j-1} - A_j D_{j-1},  D_j = sigma~_j - alpha~_j = S_j P_{j-1} - A_j D_{j-1},
    // which has two on its critical path (q = -A_j D_{j-1}, then one fma each) and, for an M-matrix, still only adds
    // terms of one sign.  sigma~, alpha~, delta~, g~ hang off P_{j-1} and are not on the chain.
    double D = sig - alp;
    double Pm = D - g;  // P_1
    double rP = pivot_rcp(Pm);
    if (KEEP) A[1] = alp, C[1] = g, S[1] = rP, F[1] = dlt;
    double G = -C0;
    s_s = fma(G * rP, sig, S0), s_f = fma(G * rP, dlt, F0);
    double lastC = g, Cprev = g;
#pragma unroll
    for (int j = 2; j < BS_R; ++j) {
      if (FULL || j < len) {
        const double nA = -(A[j] * sc), Cj = C[j] * sc, Sj = S[j] * sc;
        const double q = nA * D;
        const double Pn = fma(Sj - Cj, Pm, q);
        D = fma(Sj, Pm, q);
        sig = fma(Sj, Pm, nA * sig);
        dlt = fma(F[j] * sc, Pm, nA * dlt);
        alp = nA * alp;
        g = Cj * Pm;
        if (FULL ? j < BS_R - 1 : j < len - 1) {
          Pm = Pn;
          const double rPn = pivot_rcp(Pm);
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
  o[0] = A0, o[1] = s_c, o[2] = s_s, o[3] = s_f;
  o[4] = alp * rn, o[5] = g * rn, o[6] = sig * rn, o[7] = dlt * rn;
}
template <bool KEEP, bool FULL>
__device__ __forceinline__ void chunk0_forward(double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R], double (&F)[BS_R],
                                               int len, const LevelPtr& l1, int i0, int i1) {
  double o[8];
  chunk0_sweep<KEEP, FULL>(A, C, S, F, len, o);
  l1.a[i0] = o[0], l1.c[i0] = o[1], l1.s[i0] = o[2], l1.d[i0] = o[3];
  l1.a[i1] = o[4], l1.c[i1] = o[5], l1.s[i1] = o[6], l1.d[i1] = o[7];
}
// Back-substitution of a chunk whose sweep kept its multipliers: u of its first / last row known, u of all its rows out.
__device__ __forceinline__ void chunk0_backward(const double (&A)[BS_R], const double (&C)[BS_R], const double (&S)[BS_R],
                                                const double (&F)[BS_R], int len, double us, double ue, double (&o)[BS_R]) {
  if (len == BS_R) {
    o[0] = us, o[BS_R - 1] = ue;
#pragma unroll
    for (int j = BS_R - 2; j >= 1; --j) o[j] = fma(-C[j], o[j + 1], fma(-A[j], us, F[j])) * S[j];
  } else {
    double nxt = ue;
#pragma unroll
    for (int j = BS_R - 1; j >= 0; --j) {
      if (j == len - 1) o[j] = ue;
      else if (j == 0) o[j] = us;
      else if (j < len - 1) nxt = fma(-C[j], nxt, fma(-A[j], us, F[j])) * S[j], o[j] = nxt;
      else o[j] = 0.0;
    }
  }
}

// ---- tile kernels -------------------------------------------------------------------------------------------------
// SOLVE == false: forward sweeps, the tile's two reduced rows -> red_{a,c,s,f}[2 tile + {0,1}].
// SOLVE == true : forward sweeps again, (u_s, u_e) of the tile from ured[2 tile + {0,1}], back-substitution, out[]
//                 (ACCUM: out[] += value, the iterative-refinement update).
// SRC: p0..p3 = (lo, di, up, f) of the caller's system, or (a, c, s, f) of a reduced system.
// NLEV: shared-memory levels the tile goes through (0: all, the tile leaves 2 rows; 1: it leaves the 128 rows of level 2).
//
// Two front ends share tile_body(): k_tri_tile loads its chunk straight into registers (any alignment, any size), and
// k_tri_tile_tma (below) is the persistent, TMA-staged version used for the big level-0 systems.
struct TileArgs {
  const double *p0, *p1, *p2, *p3;
  long long m;
  double *red_a, *red_c, *red_s, *red_f;
  const double* ured;
  double* out;
  int flags;
  double bc0 = 0.0, bc1 = 0.0;  // BS_SRC_NODES: the Dirichlet values (the bands are never materialised)
};
struct ChunkGeom {
  int len;       // rows of the chunk (0: idle thread)
  long long g0;  // its first row
};
__device__ __forceinline__ ChunkGeom chunk_geom(int rows0, int c, long long tile0) {
  ChunkGeom g;
  g.len = chunk_len(rows0, c);
  const bool shifted = (rows0 & 7) == 1 && rows0 > 1 && c == ((rows0 - 1) >> 3);  // the "2" of a 7 + 2 split
  g.g0 = tile0 + (long long)c * BS_R - (shifted ? 1 : 0);
  return g;
}
// first row and row count of tile `tile` of an m-row system cut into TILE-row tiles; a single left-over row is
// never left to the last tile: the last two tiles split (TILE - 1) + 2 instead
__device__ __forceinline__ int tile_rows(long long tile, long long m, int TILE, long long* tile0) {
  const long long ntiles = (m + TILE - 1) / TILE;
  *tile0 = tile * TILE;
  int rows0 = (int)min((long long)TILE, m - *tile0);
  if (m - (ntiles - 1) * TILE == 1 && ntiles > 1) {
    if (tile == ntiles - 1) *tile0 -= 1, rows0 = 2;
    else if (tile == ntiles - 2) rows0 = TILE - 1;
  }
  return rows0;
}
// rows g0 .. g0+len of the system -> registers, straight from global memory; returns whether the 128-bit path was taken
template <int SRC, bool SOLVE>
__device__ __forceinline__ bool chunk_load_direct(const TileArgs& k, long long g0, int len, double (&A)[BS_R], double (&C)[BS_R],
                                                  double (&S)[BS_R], double (&F)[BS_R]) {
  const bool vec = len == BS_R &&  // full chunk whose 128-bit accesses are aligned
                   !((reinterpret_cast<size_t>(k.p0 + g0) | reinterpret_cast<size_t>(k.p1 + g0) | reinterpret_cast<size_t>(k.p2 + g0) |
                      reinterpret_cast<size_t>(k.p3 + g0) | (SOLVE ? reinterpret_cast<size_t>(k.out + g0) : 0)) & 15);
  if (vec) {
    const double2* q0 = reinterpret_cast<const double2*>(k.p0 + g0);
    const double2* q1 = reinterpret_cast<const double2*>(k.p1 + g0);
    const double2* q2 = reinterpret_cast<const double2*>(k.p2 + g0);
    const double2* q3 = reinterpret_cast<const double2*>(k.p3 + g0);
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
      if (SRC == BS_SRC_BANDS) A[j] = in ? k.p0[g] : 0.0, S[j] = in ? k.p1[g] : 1.0, C[j] = in ? k.p2[g] : 0.0;
      else A[j] = in ? k.p0[g] : 0.0, C[j] = in ? k.p1[g] : 0.0, S[j] = in ? k.p2[g] : 1.0;
      F[j] = in ? k.p3[g] : 0.0;
    }
  }
  return vec;
}

// Everything after the chunk is in registers, in two halves.  tile_forward: (a, c, s, f) form, level 0 and the
// shared-memory pyramid; returns the level the pyramid stopped at (see Pyramid::Left).  tile_backward: that level's d[] now
// holds its solved unknowns; back-substitution through the pyramid and the chunk, out[].  Every thread of the tile must
// call both.  vec: out[] takes 128-bit stores.  tile_body strings them together for the two-pass kernels.
template <int SRC, bool KEEP, int NLEV, int P1>
__device__ __forceinline__ typename Pyramid<P1>::Left tile_forward(Pyramid<P1>& pyr, const TileArgs& k, double (&A)[BS_R],
                                                                         double (&C)[BS_R], double (&S)[BS_R], double (&F)[BS_R],
                                                                         int len, long long g0, int t, int rows1) {
  const long long m = k.m;
  const int flags = k.flags;
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
  // level 0 in registers; KEEP leaves (alpha~_j, g~_j, 1/P_j, delta~_j) in (A, C, S, F)[1..len-2]
  if (len >= 2) {
    LevelPtr l1 = pyr.level(0);
    const int i0 = lvl_pos(2 * t, lvl_stride(P1)), i1 = lvl_pos(2 * t + 1, lvl_stride(P1));
    if (len == BS_R) chunk0_forward<KEEP, true>(A, C, S, F, len, l1, i0, i1);
    else chunk0_forward<KEEP, false>(A, C, S, F, len, l1, i0, i1);
  }
  // NLEV == 0: the pyramid runs down to the tile's final 2 rows.  NLEV > 0: it stops after NLEV levels and the tile
  // hands all rows of that level to the next application (2 P1 of them after one level: a reduction by 16 instead of
  // 1024).  The upper levels keep one warp busy and seven parked, so a throughput-bound first application is better off
  // leaving them to the next one, which then works on a 16x smaller system with full CTAs again.
  return pyr.template reduce_to<NLEV, KEEP>(t, rows1);
}

template <int SRC, bool ACCUM, int NLEV, int P1>
__device__ __forceinline__ void tile_backward(Pyramid<P1>& pyr, const TileArgs& k, const double (&A)[BS_R],
                                              const double (&C)[BS_R], const double (&S)[BS_R], const double (&F)[BS_R], int len,
                                              long long g0, bool vec, int t, int rows1) {
  const long long m = k.m;
  const int flags = k.flags;
  pyr.template backsolve_from<NLEV>(t, rows1);
  if (len >= 2) {
    LevelPtr l1 = pyr.level(0);
    const double us = l1.d[lvl_pos(2 * t, lvl_stride(P1))], ue = l1.d[lvl_pos(2 * t + 1, lvl_stride(P1))];
    double* __restrict__ out = k.out;
    double o[BS_R];
    // A decoupled end row of the caller's system (a Dirichlet row: lo = up = 0) is returned exactly as Thomas returns it,
    // f / di (matrix_solver/mod.rs:36-39 with a zero coupling); the scaled sweeps would be off by an ulp.
    // (BS_SRC_NODES: the end rows of a closed end ARE the Dirichlet rows (0, 1, 0, bc), whose Thomas value is bc / 1)
    const bool fix_first = g0 == 0 && !(flags & BS_OPEN_LEFT) && (SRC == BS_SRC_NODES || (SRC == BS_SRC_BANDS && k.p2[0] == 0.0));
    const bool fix_last = g0 + len == m && !(flags & BS_OPEN_RIGHT) && (SRC == BS_SRC_NODES || (SRC == BS_SRC_BANDS && k.p0[m - 1] == 0.0));
    const double v_first = !fix_first ? 0.0 : (SRC == BS_SRC_NODES ? k.bc0 / 1.0 : k.p3[0] / k.p1[0]);
    const double v_last = !fix_last ? 0.0 : (SRC == BS_SRC_NODES ? k.bc1 / 1.0 : k.p3[m - 1] / k.p1[m - 1]);
    if (vec) {
      o[0] = us, o[BS_R - 1] = ue;
#pragma unroll
      for (int j = BS_R - 2; j >= 1; --j) o[j] = fma(-C[j], o[j + 1], fma(-A[j], us, F[j])) * S[j];
      if (fix_first) o[0] = v_first;
      if (fix_last) o[BS_R - 1] = v_last;
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
      if (fix_first) o[0] = v_first;
#pragma unroll
      for (int j = 0; j < BS_R; ++j) {
        if (fix_last && j == len - 1) o[j] = v_last;
        if (j < len) out[g0 + j] = ACCUM ? out[g0 + j] + o[j] : o[j];
      }
    }
  }
}

template <int SRC, bool SOLVE, bool ACCUM, int NLEV, int P1>
__device__ __forceinline__ void tile_body(Pyramid<P1>& pyr, const TileArgs& k, double (&A)[BS_R], double (&C)[BS_R],
                                          double (&S)[BS_R], double (&F)[BS_R], int len, long long g0, bool vec, int t,
                                          long long tile, int rows1) {
  constexpr int ROUT = bs_rows_left(P1, NLEV);  // rows a full tile leaves
  static_assert(ROUT <= BS_THREADS, "one unknown per thread");
  const long long obase = (long long)ROUT * tile;
  // SOLVE: the solved unknowns of the level the pyramid stops at are fetched now (one per thread) and land while the
  // sweeps run; read after the forward sweeps they would put an L2 round trip on every tile's critical path
  double u_left = 0.0;
  if (SOLVE && t < bs_rows_left_of(rows1, P1, NLEV)) u_left = k.ured[obase + t];
  const typename Pyramid<P1>::Left left = tile_forward<SRC, SOLVE, NLEV>(pyr, k, A, C, S, F, len, g0, t, rows1);
  if (!SOLVE) {
    for (int r = t; r < left.rows; r += BS_THREADS) {
      const int q = left.pos(r);
      k.red_a[obase + r] = pyr.a[q], k.red_c[obase + r] = pyr.c[q], k.red_s[obase + r] = pyr.s[q], k.red_f[obase + r] = pyr.d[q];
    }
  } else {
    if (t < left.rows) pyr.d[left.pos(t)] = u_left;
    tile_backward<SRC, ACCUM, NLEV>(pyr, k, A, C, S, F, len, g0, vec, t, rows1);
  }
}

template <int SRC, bool SOLVE, bool ACCUM, int NLEV>
__global__ void __launch_bounds__(BS_THREADS, BS_MIN_CTAS) k_tri_tile(const TileArgs k) {
  __shared__ Pyramid<BS_THREADS / 4> pyr;
  const int t = threadIdx.x;
  const long long tile = blockIdx.x;
  long long tile0;
  const int rows0 = tile_rows(tile, k.m, BS_TILE, &tile0);        // rows of this tile
  const int rows1 = 2 * ((rows0 + BS_R - 1) / BS_R);              // rows it leaves at level 1
  const ChunkGeom gm = chunk_geom(rows0, t, tile0);
  double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
  const bool vec = chunk_load_direct<SRC, SOLVE>(k, gm.g0, gm.len, A, C, S, F);
  tile_body<SRC, SOLVE, ACCUM, NLEV>(pyr, k, A, C, S, F, gm.len, gm.g0, vec, t, tile, rows1);
}

// ---- TMA-staged persistent tile kernel (level 0 of a big system given as bands) ----------------------------------------
// k_tri_tile's loads are its weak spot: a thread owns 8 consecutive rows, so a warp's 128-bit load touches 16 lines (16 L1
// wavefronts per instruction; ncu: l1tex data pipe 67 % busy at 42 % DRAM), and nothing is in flight while a CTA
// computes.  Here each of the four arrays is described to the TMA as a 2-D tensor of 128-byte rows (16 doubles); one
// elected thread asks for the tile's 128 x 128 B box of each array (4 x 16 KB, SWIZZLE_128B) and the CTA's threads then
// read their 64-byte chunks from shared memory conflict-free (the swizzle XORs the 16-byte column with the row, which
// spreads the 8 threads of an LDS.128 phase over all 32 banks).  CTAs are persistent (2 per SM, tiles strided by the
// grid): as soon as every thread has its chunk in registers the stage buffer is re-armed for the CTA's NEXT tile, so that
// tile's HBM latency is hidden behind this tile's sweeps and pyramid.  Only full, 16-byte aligned tiles take the staged
// path; a ragged last tile goes through chunk_load_direct.
constexpr int BS_TMA_ROW = 16;                                      // doubles per 128-byte tensor row
constexpr int BS_TMA_BOX = BS_TILE / BS_TMA_ROW;                    // tensor rows per tile
constexpr unsigned BS_STAGE_BYTES = 4u * BS_TILE * sizeof(double);  // 64 KB
struct TileStage {
  alignas(1024) double rows[4][BS_TILE];
  Pyramid<BS_THREADS / 4> pyr;
  alignas(8) unsigned long long mbar;
};
constexpr size_t BS_STAGE_SMEM = sizeof(TileStage);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  long long t0 = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    if (!done && (spin & 1023u) == 1023u) {  // a lost copy becomes an error after ~2 s, not a hang
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, int row, unsigned long long* b) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(0), "r"(row), "r"(smem_u32(b))
      : "memory");
}

template <int SRC, bool SOLVE, bool ACCUM, int NLEV>
__global__ void __launch_bounds__(BS_THREADS, 2) k_tri_tile_tma(const __grid_constant__ CUtensorMap map0,
                                                                const __grid_constant__ CUtensorMap map1,
                                                                const __grid_constant__ CUtensorMap map2,
                                                                const __grid_constant__ CUtensorMap map3, const TileArgs k) {
  extern __shared__ __align__(1024) unsigned char tile_stage_raw[];  // the kernel has no static shared memory
  TileStage& sm = *reinterpret_cast<TileStage*>(tile_stage_raw);
  if (smem_u32(tile_stage_raw) & 1023) __trap();  // SWIZZLE_128B needs the stage on a 1024-byte boundary
  const int t = threadIdx.x;
  const long long ntiles = (k.m + BS_TILE - 1) / BS_TILE;
  if (t == 0) mbar_init(&sm.mbar, 1);
  __syncthreads();
  auto request = [&](long long tile) {  // thread 0: stage the four 16 KB boxes of a full tile
    // the stage was just read with ordinary loads (ordered before this point by the CTA barrier) and is about to be written by
    // the async proxy: the issuing thread fences the two proxies (see onelaunch_kernels.cuh, request_x, for what happens without)
    asm volatile("fence.proxy.async;" ::: "memory");
    mbar_arrive_expect_tx(&sm.mbar, BS_STAGE_BYTES);
    const int row = (int)(tile * BS_TMA_BOX);
    tma_load_box(sm.rows[0], &map0, row, &sm.mbar);
    tma_load_box(sm.rows[1], &map1, row, &sm.mbar);
    tma_load_box(sm.rows[2], &map2, row, &sm.mbar);
    tma_load_box(sm.rows[3], &map3, row, &sm.mbar);
  };
  auto staged = [&](long long tile) {
    long long t0;
    return tile < ntiles && tile_rows(tile, k.m, BS_TILE, &t0) == BS_TILE;
  };
  long long tile = blockIdx.x;
  if (t == 0 && staged(tile)) request(tile);
  unsigned phase = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    long long tile0;
    const int rows0 = tile_rows(tile, k.m, BS_TILE, &tile0);
    const int rows1 = 2 * ((rows0 + BS_R - 1) / BS_R);
    const ChunkGeom gm = chunk_geom(rows0, t, tile0);
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
    bool vec;
    if (rows0 == BS_TILE) {
      mbar_wait(&sm.mbar, phase);
      phase ^= 1;
      // chunk t = bytes [64 t, 64 t + 64) of every array = 16-byte columns 4 (t & 1) + j of tensor row t / 2
      const int row = t >> 1;
      const unsigned char* base = reinterpret_cast<const unsigned char*>(sm.rows[0]) + row * 128;
#pragma unroll
      for (int j = 0; j < BS_R / 2; ++j) {
        const int col = ((((t & 1) << 2) + j) ^ (row & 7)) << 4;
        const double2 v0 = *reinterpret_cast<const double2*>(base + col);
        const double2 v1 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 8);
        const double2 v2 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 16);
        const double2 v3 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 24);
        if (SRC == BS_SRC_BANDS) {
          A[2 * j] = v0.x, A[2 * j + 1] = v0.y, S[2 * j] = v1.x, S[2 * j + 1] = v1.y;  // S holds di for now
          C[2 * j] = v2.x, C[2 * j + 1] = v2.y;
        } else {
          A[2 * j] = v0.x, A[2 * j + 1] = v0.y, C[2 * j] = v1.x, C[2 * j + 1] = v1.y;
          S[2 * j] = v2.x, S[2 * j + 1] = v2.y;
        }
        F[2 * j] = v3.x, F[2 * j + 1] = v3.y;
      }
      __syncthreads();  // every chunk is in registers (and the previous tile is done with the pyramid)
      if (t == 0 && staged(tile + gridDim.x)) request(tile + gridDim.x);
      vec = !SOLVE || !(reinterpret_cast<size_t>(k.out + gm.g0) & 15);
    } else {
      __syncthreads();  // the previous tile is done with the pyramid
      vec = chunk_load_direct<SRC, SOLVE>(k, gm.g0, gm.len, A, C, S, F);
    }
    tile_body<SRC, SOLVE, ACCUM, NLEV>(sm.pyr, k, A, C, S, F, gm.len, gm.g0, vec, t, tile, rows1);
  }
}

// ---- single-CTA kernel: solves an m-row system (m <= 8 P1 <= 2048) entirely in shared memory ------------------------
// The rows are loaded coalesced straight into level 1 of the pyramid; inputs are not modified.  P1 in {4, 16, 64, 256}: 32 / 128 / 512 / 2048 rows.
// MODE: BS_MID_SOLVE      closed system -> out[0..m)
//       BS_MID_REDUCE     open block: the two rows left after eliminating everything inside -> top8 = (a,c,s,f) x 2
//       BS_MID_BACKSOLVE  open block whose end unknowns (us, ue) are now known -> out[0..m)
template <int P1, int SRC, int MODE, bool ACCUM>
__device__ __forceinline__ void mid_body(Pyramid<P1>& pyr, const double* p0, const double* p1, const double* p2, const double* p3,
                                         int m, double* out, int flags, double* top8, double us, double ue) {
  const int t = threadIdx.x;
  const bool open_left = flags & BS_OPEN_LEFT, open_right = flags & BS_OPEN_RIGHT;
  LevelPtr l1 = pyr.level(0);
  for (int r = t; r < m; r += BS_THREADS) {
    const int p = row_slot(r, m, lvl_stride(P1));
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
    double v = l1.d[row_slot(r, m, lvl_stride(P1))];
    if (SRC == BS_SRC_BANDS) {  // exact Dirichlet end rows, as in the tile kernels
      if (r == 0 && !open_left && p2[0] == 0.0) v = p3[0] / p1[0];
      if (r == m - 1 && !open_right && p0[m - 1] == 0.0) v = p3[m - 1] / p1[m - 1];
    }
    out[r] = ACCUM ? out[r] + v : v;
  }
}
template <int P1, int SRC, int MODE, bool ACCUM>
__global__ void __launch_bounds__(BS_THREADS) k_tri_mid(const double* __restrict__ p0, const double* __restrict__ p1,
                                                        const double* __restrict__ p2, const double* __restrict__ p3, int m,
                                                        double* __restrict__ out, int flags, double* __restrict__ top8,
                                                        double us, double ue) {
  extern __shared__ double4 mid_smem[];
  mid_body<P1, SRC, MODE, ACCUM>(*reinterpret_cast<Pyramid<P1>*>(mid_smem), p0, p1, p2, p3, m, out, flags, top8, us, ue);
}

// ---- one launch for a closed system of up to a few hundred tiles -----------------------------------------------------------
// Up to a few hundred thousand rows a solve is three dependent launches (reduce, k_tri_mid, solve) of a few microseconds
// each, and the third one reads the system and redoes the forward sweeps only because the first could not wait for the
// second.  In a cooperative launch it can: every CTA keeps its tile's multipliers in registers and shared memory across
// two grid-wide barriers, between which CTA 0 solves the 2-rows-per-tile system.  Used for a whole solve of 2049 .. ~900 000
// rows (SRC = bands: 22.7 -> 12-17 us), and for the LAST grid level of a bigger one (SRC = rows: the 625 000-row system a
// 10^7-row solve leaves after its shallow first application).  MAXT: tiles the in-kernel middle solve has room for
// (256: two CTAs per SM with all registers; 512: three, up to the co-residency limit of 444 on a B200).
template <int MAXT>
struct FusedSmem {
  Pyramid<BS_THREADS / 4> pyr;
  Pyramid<MAXT / 4> mid;  // 2 rows per tile = 8 rows per MAXT / 4 chunks
};
template <int SRC, bool ACCUM, int MAXT>
__global__ void __launch_bounds__(BS_THREADS, MAXT > 256 ? 3 : 2) k_tri_fused(const TileArgs k) {
  extern __shared__ double4 fused_smem[];
  FusedSmem<MAXT>& sm = *reinterpret_cast<FusedSmem<MAXT>*>(fused_smem);
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int t = threadIdx.x;
  const long long tile = blockIdx.x;
  long long tile0;
  const int rows0 = tile_rows(tile, k.m, BS_TILE, &tile0);
  const int rows1 = 2 * ((rows0 + BS_R - 1) / BS_R);
  const ChunkGeom gm = chunk_geom(rows0, t, tile0);
  double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
  const bool vec = chunk_load_direct<SRC, true>(k, gm.g0, gm.len, A, C, S, F);
  const Pyramid<BS_THREADS / 4>::Left left = tile_forward<SRC, true, 0>(sm.pyr, k, A, C, S, F, gm.len, gm.g0, t, rows1);
  if (t < 2) {
    const int q = left.pos(t);
    k.red_a[2 * tile + t] = sm.pyr.a[q], k.red_c[2 * tile + t] = sm.pyr.c[q], k.red_s[2 * tile + t] = sm.pyr.s[q],
                       k.red_f[2 * tile + t] = sm.pyr.d[q];
  }
  grid.sync();
  if (blockIdx.x == 0)
    mid_body<MAXT / 4, BS_SRC_ROWS, BS_MID_SOLVE, false>(sm.mid, k.red_a, k.red_c, k.red_s, k.red_f, 2 * (int)gridDim.x,
                                                        const_cast<double*>(k.ured), 0, nullptr, 0.0, 0.0);
  grid.sync();
  if (t < 2) sm.pyr.d[left.pos(t)] = __ldcg(k.ured + 2 * tile + t);
  tile_backward<SRC, ACCUM, 0>(sm.pyr, k, A, C, S, F, gm.len, gm.g0, vec, t, rows1);
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

// (the SPIKE interface system of a bar split across GPUs: comm_kernels.cuh)

// ---- is the system one the partitioned elimination is reliable on? -------------------------------------------------
// The chunk sweeps eliminate every chunk on its own (local Dirichlet problems).  That is as stable as serial Thomas
// for an M-matrix (lo, up <= 0 and the row sum not below -1e-6 of the diagonal: weak dominance up to assembly noise) and
// for a strictly dominant matrix with positive couplings (the mass matrix); on an indefinite system (e.g. the reference's
// kink-straddling quadrature on a strongly irregular mesh, or a cell Peclet number above 1) the local problems can be
// nearly singular while the serial recurrence, which always carries the left boundary along, is not.
// flags: bit 0 set = not M-like, bit 1 set = not mass-like.  Interior rows only: rows 1 .. n-2 of a closed system; at an open
// end (a block of a longer bar, BS_OPEN_LEFT / BS_OPEN_RIGHT) the block's first / last row is an interior row of the bar too.
__global__ void __launch_bounds__(256) k_tri_classify(const double* __restrict__ lo, const double* __restrict__ di,
                                                      const double* __restrict__ up, long long n, int* __restrict__ flags,
                                                      int open_ends) {
  int f = 0;
  const long long first = (open_ends & BS_OPEN_LEFT) ? 0 : 1, last = (open_ends & BS_OPEN_RIGHT) ? n : n - 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + first; i < last; i += (long long)gridDim.x * blockDim.x) {
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
  int nlev = 0;  // shared-memory levels its tiles go through: 0 = all (2 rows per tile), 1 = one (128 rows per tile)
  double *a = nullptr, *c = nullptr, *s = nullptr, *f = nullptr, *u = nullptr;
};
struct BigSys {
  size_t n = 0;
  size_t shallow_min_tiles = BS_SHALLOW_MIN_TILES;  // applications with at least this many tiles stop after one level
  bool tma = true;    // big systems go through k_tri_tile_tma (DZB_NO_TMA=1: always k_tri_tile)
  bool fused = true;  // closed systems of <= BS_FUSED_MAX_TILES tiles: one cooperative launch (DZB_NO_FUSED=1: three launches)
  size_t tma_min_tiles = BS_TMA_MIN_TILES;
  struct Map {
    const double* p;
    size_t m;
    CUtensorMap map;
  };
  std::vector<Map> maps;         // tensor maps of the arrays seen so far (encoding one costs a few microseconds)
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

// One-off per-DEVICE set-up at a call site (function attributes and occupancy are per device; the ABI has dzb_set_device):
//   static PerDevice<int> cap;  int& v = cap.slot(&fresh);  if (fresh) v = ...;
constexpr int BS_MAX_DEVICES = 64;
inline int bigsys_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < BS_MAX_DEVICES ? d : 0;
}
template <class T>
struct PerDevice {
  T v[BS_MAX_DEVICES] = {};
  bool init[BS_MAX_DEVICES] = {};
  T& slot(bool* fresh) {
    const int d = bigsys_device();
    *fresh = !init[d];
    init[d] = true;
    return v[d];
  }
  void forget() { init[bigsys_device()] = false; }  // the set-up failed: try again next time
};

inline void bigsys_free(BigSys* s) {
  for (BigLevel& l : s->levels)
    for (double** p : {&l.a, &l.c, &l.s, &l.f, &l.u})
      if (*p) cudaFree(*p), *p = nullptr;
  s->levels.clear();
  if (s->resid) cudaFree(s->resid), s->resid = nullptr;
  s->n = 0;
}

// rows an m-row system leaves after one application of the tile kernel (same tiling rules as tile_rows / chunk_len)
inline size_t bigsys_rows_left(size_t m, int nlev) {
  const size_t ntiles = (m + BS_TILE - 1) / BS_TILE;
  size_t last = m - (ntiles - 1) * BS_TILE;  // rows of the last tile
  if (last == 1 && ntiles > 1) last = 2;     // the (TILE - 1) + 2 split
  size_t left = 2 * ((last + BS_R - 1) / BS_R);             // after level 0
  left = nlev ? 2 * ((left + BS_R - 1) / BS_R) : 2;         // after one level / after all of them
  return (ntiles - 1) * (size_t)bs_rows_left(BS_THREADS / 4, nlev) + left;
}

// Allocates the scratch pyramid for systems of n rows.
inline int bigsys_prepare(BigSys* s, size_t n) {
  bigsys_free(s);
  s->n = n;
  if (const char* v = std::getenv("DZB_SHALLOW_MIN_TILES")) s->shallow_min_tiles = (size_t)std::atoll(v);  // tuning knob
  if (const char* v = std::getenv("DZB_NO_TMA")) s->tma = std::atoi(v) == 0;
  if (const char* v = std::getenv("DZB_NO_FUSED")) s->fused = std::atoi(v) == 0;
  if (const char* v = std::getenv("DZB_TMA_MIN_TILES")) s->tma_min_tiles = (size_t)std::atoll(v);  // tuning knob
  s->maps.clear();
  cudaError_t e;
  size_t m = n;
  while (m > (size_t)BS_MID_CAP) {
    BigLevel l;
    l.nlev = (m + BS_TILE - 1) / BS_TILE >= s->shallow_min_tiles ? 1 : 0;
    l.m = bigsys_rows_left(m, l.nlev);
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
    static PerDevice<char> attr;
    bool fresh;
    attr.slot(&fresh);
    if (fresh) {
      cudaError_t e1 = cudaFuncSetAttribute(k_tri_mid<P1, SRC, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaError_t e2 = cudaFuncSetAttribute(k_tri_mid<P1, SRC, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess) {
        attr.forget();
        return bigsys_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute");
      }
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

// ---- tensor maps for the staged kernel ------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn bigsys_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// the array p[0..m) as 128-byte rows (only the floor(m / 16) complete ones: full tiles never reach past them)
inline bool bigsys_map(BigSys* s, const double* p, size_t m, CUtensorMap* out) {
  for (const BigSys::Map& e : s->maps)
    if (e.p == p && e.m == m) {
      *out = e.map;
      return true;
    }
  EncodeTiledFn enc = bigsys_encoder();
  if (!enc) return false;
  BigSys::Map e;
  e.p = p, e.m = m;
  const cuuint64_t dims[2] = {(cuuint64_t)BS_TMA_ROW, (cuuint64_t)(m / BS_TMA_ROW)};
  const cuuint64_t strides[1] = {(cuuint64_t)BS_TMA_ROW * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)BS_TMA_ROW, (cuuint32_t)BS_TMA_BOX};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (s->maps.size() >= 32) s->maps.erase(s->maps.begin());
  s->maps.push_back(e);
  *out = e.map;
  return true;
}
inline int bigsys_sm_count() {
  static PerDevice<int> n;
  bool fresh;
  int& v = n.slot(&fresh);
  if (fresh) {
    int dev = 0, q = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&q, cudaDevAttrMultiProcessorCount, dev);
    v = q > 0 ? q : 148;
  }
  return v;
}

// One application of the tile kernel.  Big systems whose arrays are 16-byte aligned go through the TMA-staged persistent
// kernel; everything else (offset views of a SPIKE block, small systems) through k_tri_tile.
template <int SRC, bool SOLVE, bool ACCUM, int NLEV>
inline cudaError_t bigsys_launch_tile_n(BigSys* s, size_t m, const double* p0, const double* p1, const double* p2, const double* p3,
                                        double* ra, double* rc, double* rs, double* rf, const double* ured, double* out, int flags,
                                        cudaStream_t st) {
  const TileArgs k{p0, p1, p2, p3, (long long)m, ra, rc, rs, rf, ured, out, flags};
  const size_t tiles = (m + BS_TILE - 1) / BS_TILE;
  if (s->tma && tiles >= s->tma_min_tiles &&
      !((reinterpret_cast<size_t>(p0) | reinterpret_cast<size_t>(p1) | reinterpret_cast<size_t>(p2) | reinterpret_cast<size_t>(p3)) & 15)) {
    CUtensorMap c0, c1, c2, c3;  // by value: a lookup may grow the cache
    if (bigsys_map(s, p0, m, &c0) && bigsys_map(s, p1, m, &c1) && bigsys_map(s, p2, m, &c2) && bigsys_map(s, p3, m, &c3)) {
      static PerDevice<char> attr;  // per instantiation and device
      bool fresh;
      attr.slot(&fresh);
      if (fresh) {
        cudaError_t e = cudaFuncSetAttribute(k_tri_tile_tma<SRC, SOLVE, ACCUM, NLEV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)BS_STAGE_SMEM);
        if (e != cudaSuccess) {
          attr.forget();
          return e;
        }
      }
      const size_t grid = std::min(tiles, (size_t)2 * bigsys_sm_count());
      k_tri_tile_tma<SRC, SOLVE, ACCUM, NLEV><<<(unsigned)grid, BS_THREADS, BS_STAGE_SMEM, st>>>(c0, c1, c2, c3, k);
      return cudaGetLastError();
    }
  }
  k_tri_tile<SRC, SOLVE, ACCUM, NLEV><<<(unsigned)tiles, BS_THREADS, 0, st>>>(k);
  return cudaGetLastError();
}
template <int SRC, bool SOLVE, bool ACCUM>
inline cudaError_t bigsys_launch_tile(BigSys* s, int nlev, size_t m, const double* p0, const double* p1, const double* p2,
                                      const double* p3, double* ra, double* rc, double* rs, double* rf, const double* ured,
                                      double* out, int flags, cudaStream_t st) {
  if (nlev) return bigsys_launch_tile_n<SRC, SOLVE, ACCUM, 1>(s, m, p0, p1, p2, p3, ra, rc, rs, rf, ured, out, flags, st);
  return bigsys_launch_tile_n<SRC, SOLVE, ACCUM, 0>(s, m, p0, p1, p2, p3, ra, rc, rs, rf, ured, out, flags, st);
}

// forward sweeps of every grid level: each application of k_tri_tile leaves 2 rows per tile
inline int bigsys_forward(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                          cudaStream_t st, unsigned long long* launches, size_t kend = (size_t)-1) {  // levels [0, kend)
  cudaError_t e;
  for (size_t k = 0; k < std::min(kend, s->levels.size()); ++k) {
    BigLevel& L = s->levels[k];
    if (k == 0) {
      e = bigsys_launch_tile<BS_SRC_BANDS, false, false>(s, L.nlev, s->n, lo, di, up, f, L.a, L.c, L.s, L.f, nullptr, nullptr, flags, st);
    } else {
      BigLevel& Q = s->levels[k - 1];
      e = bigsys_launch_tile<BS_SRC_ROWS, false, false>(s, L.nlev, Q.m, Q.a, Q.c, Q.s, Q.f, L.a, L.c, L.s, L.f, nullptr, nullptr, flags, st);
    }
    ++*launches;
    if (e != cudaSuccess) return bigsys_fail(e, "k_tri_tile<reduce>");
  }
  return 0;
}
// back-substitution of every grid level; levels.back().u must hold the solved top-level unknowns
inline int bigsys_backward(BigSys* s, const double* lo, const double* di, const double* up, const double* f, int flags,
                           double* out, bool accum, cudaStream_t st, unsigned long long* launches,
                           size_t kend = (size_t)-1) {  // levels [0, kend), from the top
  cudaError_t e;
  for (size_t k = std::min(kend, s->levels.size()); k-- > 0;) {
    BigLevel& L = s->levels[k];
    if (k == 0) {
      if (accum)
        e = bigsys_launch_tile<BS_SRC_BANDS, true, true>(s, L.nlev, s->n, lo, di, up, f, nullptr, nullptr, nullptr, nullptr, L.u, out, flags, st);
      else
        e = bigsys_launch_tile<BS_SRC_BANDS, true, false>(s, L.nlev, s->n, lo, di, up, f, nullptr, nullptr, nullptr, nullptr, L.u, out, flags, st);
    } else {
      BigLevel& Q = s->levels[k - 1];
      e = bigsys_launch_tile<BS_SRC_ROWS, true, false>(s, L.nlev, Q.m, Q.a, Q.c, Q.s, Q.f, nullptr, nullptr, nullptr, nullptr, L.u, Q.u, flags, st);
    }
    ++*launches;
    if (e != cudaSuccess) return bigsys_fail(e, "k_tri_tile<solve>");
  }
  return 0;
}

// One cooperative launch for the closed m-row system (p0..p3) whose tiles each leave 2 rows in L: possible while the tiles
// are few enough for the in-kernel middle solve and all co-resident (a grid-wide barrier needs that; the capacity of
// each instantiation is asked once).  Returns -1 when it does not apply (the caller takes the three-launch path).
template <int SRC, bool ACCUM, int MAXT>
inline int bigsys_fused_try_n(const TileArgs& k, size_t tiles, cudaStream_t st, unsigned long long* launches) {
  constexpr size_t smem = sizeof(FusedSmem<MAXT>);
  static PerDevice<int> cap;  // per instantiation and device
  bool fresh;
  int& capacity = cap.slot(&fresh);
  if (fresh) {
    int dev = 0, coop = 0, per_sm = 0;
    capacity = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (coop && cudaFuncSetAttribute(k_tri_fused<SRC, ACCUM, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tri_fused<SRC, ACCUM, MAXT>, BS_THREADS, smem) == cudaSuccess)
      capacity = per_sm * bigsys_sm_count();
  }
  if (tiles > (size_t)MAXT || tiles > (size_t)capacity) return -1;
  TileArgs args = k;
  void* params[] = {&args};
  ++*launches;
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_tri_fused<SRC, ACCUM, MAXT>, dim3((unsigned)tiles), dim3(BS_THREADS), params,
                                              smem, st);
  return e == cudaSuccess ? 0 : bigsys_fail(e, "k_tri_fused");
}
template <int SRC, bool ACCUM>
inline int bigsys_fused_try(const TileArgs& k, cudaStream_t st, unsigned long long* launches) {
  const size_t tiles = ((size_t)k.m + BS_TILE - 1) / BS_TILE;
  if (tiles <= 256) return bigsys_fused_try_n<SRC, ACCUM, 256>(k, tiles, st, launches);
  return bigsys_fused_try_n<SRC, ACCUM, BS_FUSED_MAX_TILES>(k, tiles, st, launches);
}

// out (+)= A^{-1} f for the closed n-row system (lo, di, up); inputs are not modified.
inline int bigsys_solve_once(BigSys* s, const double* lo, const double* di, const double* up, const double* f, double* out,
                             bool accum, cudaStream_t st, unsigned long long* launches) {
  if (s->levels.empty())
    return bigsys_mid<BS_SRC_BANDS, BS_MID_SOLVE>(lo, di, up, f, s->n, out, accum, 0, nullptr, 0.0, 0.0, st, launches);
  const size_t K = s->levels.size();
  BigLevel& L = s->levels.back();
  // the last grid level (full pyramid, leaves 2 rows per tile) together with the middle solve in one cooperative launch
  if (s->fused && L.nlev == 0) {
    int rc = -1;
    if (K == 1) {
      const TileArgs k{lo, di, up, f, (long long)s->n, L.a, L.c, L.s, L.f, L.u, out, 0};
      rc = accum ? bigsys_fused_try<BS_SRC_BANDS, true>(k, st, launches) : bigsys_fused_try<BS_SRC_BANDS, false>(k, st, launches);
      if (rc >= 0) return rc;
    } else {
      BigLevel& Q = s->levels[K - 2];
      const size_t tiles = (Q.m + BS_TILE - 1) / BS_TILE;
      // (checked before the forward sweeps are launched; co-residency inside.)  Only up to 256 tiles: beyond that the
      // three-launch path with its TMA-staged tiles is as fast (measured at 306 tiles, the 10^7-row case: 167 vs 162 us)
      if (tiles <= 256) {
        if (bigsys_forward(s, lo, di, up, f, 0, st, launches, K - 1)) return 1;
        const TileArgs k{Q.a, Q.c, Q.s, Q.f, (long long)Q.m, L.a, L.c, L.s, L.f, L.u, Q.u, 0};
        rc = bigsys_fused_try<BS_SRC_ROWS, false>(k, st, launches);
        if (rc > 0) return rc;
        if (rc < 0) {  // not co-resident after all: the last level the ordinary way
          cudaError_t e = bigsys_launch_tile<BS_SRC_ROWS, false, false>(s, L.nlev, Q.m, Q.a, Q.c, Q.s, Q.f, L.a, L.c, L.s, L.f, nullptr,
                                                                       nullptr, 0, st);
          ++*launches;
          if (e != cudaSuccess) return bigsys_fail(e, "k_tri_tile<reduce>");
          if (bigsys_mid<BS_SRC_ROWS, BS_MID_SOLVE>(L.a, L.c, L.s, L.f, L.m, L.u, false, 0, nullptr, 0.0, 0.0, st, launches)) return 1;
          return bigsys_backward(s, lo, di, up, f, 0, out, accum, st, launches, K);
        }
        return bigsys_backward(s, lo, di, up, f, 0, out, accum, st, launches, K - 1);
      }
    }
  }
  if (bigsys_forward(s, lo, di, up, f, 0, st, launches)) return 1;
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
