// onelaunch_kernels.cuh — assemble + solve of one long bar in ONE launch, the bands never leave the chip.
//
// The reference builds K with its quadrature loop and then runs solve_by_thomas over it (diffusion_solver/
// time_independent.rs:91-196, matrix_solver/mod.rs:17-48).  The multi-launch path of bigsys_kernels.cuh mirrors that
// split: k_assemble_fast writes four bands (32 B/row), the reduce pass reads them, two or three small launches solve the
// reduced system, the solve pass reads the bands again (112 B/row moved for 80 B/row of algorithmic traffic, five
// dependent launches).  Here the whole thing is one persistent cooperative kernel:
//
//   * every CTA owns a CONTIGUOUS range of 2048-row tiles (not a strided one), so that what its tiles leave behind is a
//     contiguous piece of the reduced system and can be reduced further inside the CTA;
//   * pass 1, per tile: the node coordinates arrive through the TMA (one 16 KB box, SWIZZLE_128B, prefetched one tile
//     ahead), every thread forms the 8 rows of its chunk in registers with the closed form of assembly_kernels.cuh
//     (same operations, bit-identical to k_assemble_fast's bands), sweeps them (level 0 in registers, two shared-memory
//     levels: 2048 -> 512 -> 128 -> 32 rows) and the 32 rows that are left join the CTA's own level (32 rows per tile);
//   * the CTA reduces that level to its 2 interface rows (shared memory, multipliers kept), writes them to global memory,
//     ONE grid-wide barrier, then every CTA solves the 2-rows-per-CTA system redundantly in shared memory (592 rows on a
//     B200: 19 KB from L2) — no second barrier, no CTA waits for another one again;
//   * back-substitution through the CTA's level, then pass 2 over the CTA's tiles in reverse order (the most recently
//     read coordinates are the likeliest to still be in L2): rows formed again, swept again with the multipliers kept,
//     back-substituted, u written.
//
// HBM traffic: x twice (the second time mostly from L2) + u once = 16-24 B/row instead of 112; the price is forming the
// rows twice, which makes the kernel fp64-issue bound rather than HBM bound.  The same kernel also takes materialised
// bands (BS_SRC_BANDS: the reference-arithmetic assembly writes them once, 72 B/row here instead of 72 B/row + four launches).
#pragma once
#include "assembly_kernels.cuh"
#include "bigsys_kernels.cuh"

namespace dzb {

constexpr int OL_NLEV = 2;                                         // shared-memory levels a tile goes through
constexpr int OL_ROUT = bs_rows_left(BS_THREADS / 4, OL_NLEV);     // rows a full tile leaves: 32
constexpr int OL_P1M = 96;                                         // middle system: 2 rows per CTA, up to 768 rows (384 CTAs)
constexpr int OL_QX_CAP = 512;                                     // quadrature nodes kept in shared memory for the kink probes
static_assert(OL_ROUT == 32, "tile -> CTA level hand-over");

struct OneLaunchArgs {
  TileArgs k;        // NODES: p0 = x with x[r] the node of row r; BANDS / ROWS: the four arrays.  red_* : [2 gridDim.x] scratch
  RuleView rule;     // NODES
  AsmParams P;       // NODES
  long long n_local; // NODES: nodes of the local coordinate array (rows + halo nodes of an open block)
  int halo_left;     // NODES: 1 when x[-1] exists (open left end): local node index of row r = r + halo_left
  int* class_flags;  // NODES: k_tri_classify's bits, or'ed over every interior row (bit 0: not M-like, bit 1: not mass-like)
  unsigned* bad_rows; // NODES: rows that needed the quadrature loop (statistics only)
};

template <int NARR, int P1C>
struct OneLaunchSmem {
  union Work {
    struct Tile {
      alignas(1024) double rows[NARR][BS_TILE];   // TMA stage
      Pyramid<BS_THREADS / 4> pyr;
    } tile;
    Pyramid<OL_P1M> mid;                          // only between the two passes, when no copy is in flight
  } w;
  Pyramid<P1C> cta;
  double qx[OL_QX_CAP];
  alignas(8) unsigned long long mbar;
};

// element e (0 .. 2047) of a staged tile: 128-byte tensor rows, 16-byte columns XOR-swizzled with the row (SWIZZLE_128B)
__device__ __forceinline__ double stage_elem(const double* stage, int e) {
  const int row = e >> 4, col = (e >> 1) & 7;
  return *reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(stage) + row * 128 + ((col ^ (row & 7)) << 4) + (e & 1) * 8);
}

__device__ __noinline__ void ol_quadrature_row(const double* xl, long long n_local, long long i, RuleView rule, AsmParams P, double* out7) {
  double v[7];
  faithful_row<KIND_TI>(xl, n_local, i, rule.x, rule.w, rule.m, P, nullptr, 0, v);
#pragma unroll
  for (int q = 0; q < 7; ++q) out7[q] = v[q];
}

// Rows i0 .. i0+7 (local node indices) of the static diffusion system in row-excess form (a, c, s, f), from the
// coordinates X[k] = x_local[i0 - 2 + k] (entries outside [0, n_local) are never used).  The arithmetic is fast_row<KIND_TI>'s
// (assembly_kernels.cuh), shared between the two rows that touch an element: bands bit-identical to k_assemble_fast's.
template <bool CLASSIFY>
__device__ __forceinline__ void ti_rows_from_nodes(const double (&X)[12], long long i0, int len, const OneLaunchArgs& a,
                                                   const double* __restrict__ qx, double (&A)[BS_R], double (&C)[BS_R],
                                                   double (&S)[BS_R], double (&F)[BS_R], int& cls, unsigned& nbad) {
  const RuleView& R = a.rule;
  const int m = R.m;
  const double x_hi = qx[0], x_lo = qx[m - 1];
  const double K0 = -0.5 * R.t0;                       // prime_{prev,next} = K0 * (1/h)
  const double bhp = a.P.b * (-0.25 * R.t1p), bhn = a.P.b * (0.25 * R.t1m);
  // per element el = 0..8 between X[el+1] and X[el+2]
  double h[BS_R + 1], r[BS_R + 1];
  bool ok_el[BS_R + 1];
#pragma unroll
  for (int el = 0; el <= BS_R; ++el) {
    const double xa = X[el + 1], xb = X[el + 2];
    h[el] = xb - xa;
    r[el] = __drcp_rn(h[el]);
    const Affine tm = from_m1_p1(xa, xb);
    ok_el[el] = (xa < xb) && (tm(x_hi) < xb) && !(tm(x_lo) < xa);
  }
#pragma unroll
  for (int j = 0; j < BS_R; ++j) {
    const long long i = i0 + j;
    double lo = 0.0, di = 1.0, up = 0.0, f = 0.0, ex = 1.0;  // padding and Dirichlet rows: (0, 1, 0, f), excess 1
    if (j < len) {
      if (i == 0) f = a.P.bc0;
      else if (i == a.n_local - 1) f = a.P.bc1;
      else {
        const double xa = X[j + 1], xc = X[j + 2], xe = X[j + 3];
        const Affine ts = from_m1_p1(xa, xe);
        bool ok = ok_el[j] && ok_el[j + 1] && (ts(x_hi) < xe) && !(ts(x_lo) < xa);
        if (i >= 2) ok = ok && !(xa < X[j]);
        if (i + 2 < a.n_local) ok = ok && !(X[j + 4] < xe);
        if (ok) {
          const double h1 = h[j], h2 = h[j + 1], H = xe - xa;
          int s;
          {  // first node whose image is left of the kink: bracket table + probes with the reference's own predicate
            const float xs = __fdividef((float)(h1 - h2), (float)H);
            int bin = (int)((xs + 1.0f) * (0.5f * GUESS_BINS));
            bin = max(0, min(GUESS_BINS - 1, bin));
            s = R.guess[bin + 1];
            while (s < m && !(ts(qx[s]) < xc)) ++s;
            while (s > 0 && (ts(qx[s - 1]) < xc)) --s;
          }
          const double r1 = r[j], r2 = r[j + 1];
          const double ih1sq = r1 * r1, ih2sq = r2 * r2;
          const double U0 = R.u0[s], U1 = R.u1[s], V0 = R.v0[s], V1 = R.v1[s];
          const double prime_sq = (0.5 * H) * fma(U0, ih1sq, V0 * ih2sq);
          const double half_sq = (0.25 * H * H) * fma(U1, ih1sq, -V1 * ih2sq);
          lo = fma(a.P.mu, K0 * r1, bhp);
          di = fma(a.P.mu, prime_sq, a.P.b * half_sq);
          up = fma(a.P.mu, K0 * r2, bhn);
        } else {  // degenerate / unsorted spacing: the reference's quadrature loop decides which piece every node hits
          double v[7];
          ol_quadrature_row(a.k.p0 - a.halo_left, a.n_local, i, R, a.P, v);
          lo = v[0], di = v[1], up = v[2], f = v[3];
          ++nbad;
        }
        ex = excess3(lo, di, up);
        if (CLASSIFY) {
          if (!(lo <= 0.0 && up <= 0.0 && ex >= -1e-6 * di)) cls |= 1;
          if (!(lo >= 0.0 && up >= 0.0 && di >= 1.25 * (lo + up))) cls |= 2;
        }
      }
    }
    A[j] = lo, C[j] = up, S[j] = ex, F[j] = f;  // row-excess form: tile_forward<BS_SRC_ROWS> takes it as it is
  }
}

// The chunk of thread t: rows g0 .. g0+len of the system.  Staged tiles hand over the 8 own coordinates from shared
// memory and the 2 + 2 halo coordinates from the neighbouring lanes; everything else reads global memory directly.
__device__ __forceinline__ void nodes_from_stage(const double* stage, const OneLaunchArgs& a, long long g0, int t, double (&X)[12]) {
  const int row = t >> 1;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(stage) + row * 128;
#pragma unroll
  for (int j = 0; j < BS_R / 2; ++j) {
    const int col = ((((t & 1) << 2) + j) ^ (row & 7)) << 4;
    const double2 v = *reinterpret_cast<const double2*>(base + col);
    X[2 + 2 * j] = v.x, X[3 + 2 * j] = v.y;
  }
  const int lane = t & 31;
  X[0] = __shfl_up_sync(0xffffffffu, X[8], 1);
  X[1] = __shfl_up_sync(0xffffffffu, X[9], 1);
  X[10] = __shfl_down_sync(0xffffffffu, X[2], 1);
  X[11] = __shfl_down_sync(0xffffffffu, X[3], 1);
  const long long i0 = g0 + a.halo_left;  // local node index of the chunk's first row
  const double* xl = a.k.p0 - a.halo_left;
  if (lane == 0) {
    if (t > 0) X[0] = stage_elem(stage, 8 * t - 2), X[1] = stage_elem(stage, 8 * t - 1);
    else X[0] = i0 >= 2 ? xl[i0 - 2] : 0.0, X[1] = i0 >= 1 ? xl[i0 - 1] : 0.0;
  }
  if (lane == 31) {
    if (t < BS_THREADS - 1) X[10] = stage_elem(stage, 8 * t + 8), X[11] = stage_elem(stage, 8 * t + 9);
    else X[10] = i0 + 8 < a.n_local ? xl[i0 + 8] : 0.0, X[11] = i0 + 9 < a.n_local ? xl[i0 + 9] : 0.0;
  }
}
__device__ __forceinline__ void nodes_direct(const OneLaunchArgs& a, long long g0, double (&X)[12]) {
  const long long i0 = g0 + a.halo_left;
  const double* xl = a.k.p0 - a.halo_left;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const long long i = i0 - 2 + q;
    X[q] = (i >= 0 && i < a.n_local) ? xl[i] : 0.0;
  }
}

template <int SRC>
struct OlSrc {  // what tile_forward / tile_backward are told: NODES hands them freshly formed rows (a, c, s, f)
  static constexpr int FWD = SRC == BS_SRC_NODES ? BS_SRC_ROWS : SRC;
  static constexpr int BWD = SRC;
  static constexpr int NARR = SRC == BS_SRC_NODES ? 1 : 4;
};

template <int SRC, bool ACCUM, int P1C>
__global__ void __launch_bounds__(BS_THREADS, 2) k_tri_onelaunch(const __grid_constant__ CUtensorMap map0,
                                                                 const __grid_constant__ CUtensorMap map1,
                                                                 const __grid_constant__ CUtensorMap map2,
                                                                 const __grid_constant__ CUtensorMap map3, const OneLaunchArgs a,
                                                                 const int use_tma) {
  using Smem = OneLaunchSmem<OlSrc<SRC>::NARR, P1C>;
  constexpr int NARR = OlSrc<SRC>::NARR;
  constexpr int P1T = BS_THREADS / 4;
  extern __shared__ __align__(1024) unsigned char ol_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(ol_smem_raw);
  if (smem_u32(ol_smem_raw) & 1023) __trap();
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const TileArgs& k = a.k;
  const int t = threadIdx.x;
  const long long ntiles = (k.m + BS_TILE - 1) / BS_TILE;
  const long long tb0 = (long long)blockIdx.x * ntiles / gridDim.x, tb1 = ((long long)blockIdx.x + 1) * ntiles / gridDim.x;
  const int T = (int)(tb1 - tb0);  // >= 1: the launcher never starts more CTAs than tiles
  int rowsC;                       // rows of this CTA's level: 32 per tile (any short tile is the last one of the system)
  {
    long long tl0;
    const int rl = tile_rows(tb1 - 1, k.m, BS_TILE, &tl0);
    rowsC = OL_ROUT * (T - 1) + bs_rows_left_of(2 * ((rl + BS_R - 1) / BS_R), P1T, OL_NLEV);
  }
  constexpr int PC = lvl_stride(P1C);
  const double* qx = a.rule.x;
  if (SRC == BS_SRC_NODES && a.rule.m <= OL_QX_CAP) {
    for (int j = t; j < a.rule.m; j += BS_THREADS) sm.qx[j] = a.rule.x[j];
    qx = sm.qx;
  }
  if (t == 0) mbar_init(&sm.mbar, 1);
  __syncthreads();
  auto staged = [&](long long tile) {
    long long t0;
    return use_tma && tile_rows(tile, k.m, BS_TILE, &t0) == BS_TILE;
  };
  auto request = [&](long long tile) {  // thread 0
    mbar_arrive_expect_tx(&sm.mbar, (unsigned)(NARR * BS_TILE * sizeof(double)));
    const int row = (int)(tile * BS_TMA_BOX);
    tma_load_box(sm.w.tile.rows[0], &map0, row, &sm.mbar);
    if constexpr (NARR == 4) {
      tma_load_box(sm.w.tile.rows[NARR - 3], &map1, row, &sm.mbar);
      tma_load_box(sm.w.tile.rows[NARR - 2], &map2, row, &sm.mbar);
      tma_load_box(sm.w.tile.rows[NARR - 1], &map3, row, &sm.mbar);
    }
  };
  unsigned phase = 0;
  int cls = 0;
  unsigned nbad = 0;
  // One tile into registers as (lo, di, up, f) / (a, c, s, f); `next` is the tile to prefetch once the stage is free.
  auto load_tile = [&](long long tile, long long next, bool classify, double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R],
                       double (&F)[BS_R], ChunkGeom& gm, int& rows1, bool& vec) {
    long long tile0;
    const int rows0 = tile_rows(tile, k.m, BS_TILE, &tile0);
    rows1 = 2 * ((rows0 + BS_R - 1) / BS_R);
    gm = chunk_geom(rows0, t, tile0);
    const bool st = staged(tile);
    if (st) {
      mbar_wait(&sm.mbar, phase);
      phase ^= 1;
    }
    if (SRC == BS_SRC_NODES) {
      double X[12];
      if (st) nodes_from_stage(sm.w.tile.rows[0], a, gm.g0, t, X);
      else nodes_direct(a, gm.g0, X);
      __syncthreads();  // every chunk is in registers, and the previous tile is done with the pyramid
      if (t == 0 && next >= 0 && staged(next)) request(next);
      const long long i0 = gm.g0 + a.halo_left;
      if (classify) ti_rows_from_nodes<true>(X, i0, gm.len, a, qx, A, C, S, F, cls, nbad);
      else ti_rows_from_nodes<false>(X, i0, gm.len, a, qx, A, C, S, F, cls, nbad);
      vec = gm.len == BS_R && !(reinterpret_cast<size_t>(k.out + gm.g0) & 15);
    } else {
      if (st) {
        const int row = t >> 1;
        const unsigned char* base = reinterpret_cast<const unsigned char*>(sm.w.tile.rows[0]) + row * 128;
#pragma unroll
        for (int j = 0; j < BS_R / 2; ++j) {
          const int col = ((((t & 1) << 2) + j) ^ (row & 7)) << 4;
          const double2 v0 = *reinterpret_cast<const double2*>(base + col);
          const double2 v1 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 8);
          const double2 v2 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 16);
          const double2 v3 = *reinterpret_cast<const double2*>(base + col + BS_TILE * 24);
          if (SRC == BS_SRC_BANDS) {
            A[2 * j] = v0.x, A[2 * j + 1] = v0.y, S[2 * j] = v1.x, S[2 * j + 1] = v1.y;
            C[2 * j] = v2.x, C[2 * j + 1] = v2.y;
          } else {
            A[2 * j] = v0.x, A[2 * j + 1] = v0.y, C[2 * j] = v1.x, C[2 * j + 1] = v1.y;
            S[2 * j] = v2.x, S[2 * j + 1] = v2.y;
          }
          F[2 * j] = v3.x, F[2 * j + 1] = v3.y;
        }
        __syncthreads();
        if (t == 0 && next >= 0 && staged(next)) request(next);
        vec = !(reinterpret_cast<size_t>(k.out + gm.g0) & 15);
      } else {
        __syncthreads();
        if (t == 0 && next >= 0 && staged(next)) request(next);
        vec = chunk_load_direct<OlSrc<SRC>::FWD, true>(k, gm.g0, gm.len, A, C, S, F);
      }
    }
  };

  // ---- pass 1: every tile down to its 32 rows, which join the CTA's level
  if (t == 0 && staged(tb0)) request(tb0);
  LevelPtr c1 = sm.cta.level(0);
  for (int kk = 0; kk < T; ++kk) {
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
    ChunkGeom gm;
    int rows1;
    bool vec;
    load_tile(tb0 + kk, kk + 1 < T ? tb0 + kk + 1 : -1, a.class_flags != nullptr, A, C, S, F, gm, rows1, vec);
    const typename Pyramid<P1T>::Left left =
        tile_forward<OlSrc<SRC>::FWD, false, OL_NLEV>(sm.w.tile.pyr, k, A, C, S, F, gm.len, gm.g0, t, rows1);
    if (t < left.rows) {
      const int q = left.pos(t), p = row_slot(OL_ROUT * kk + t, rowsC, PC);
      c1.a[p] = sm.w.tile.pyr.a[q], c1.c[p] = sm.w.tile.pyr.c[q], c1.s[p] = sm.w.tile.pyr.s[q], c1.d[p] = sm.w.tile.pyr.d[q];
    }
  }
  if (SRC == BS_SRC_NODES) {
    if (cls && a.class_flags) atomicOr(a.class_flags, cls);
    if (nbad) atomicAdd(a.bad_rows, nbad);
  }
  // ---- the CTA's level down to its two interface rows
  sm.cta.template reduce_to<0, true>(t, rowsC);
  {
    LevelPtr top = sm.cta.top();
    if (t < 2) {
      const long long o = 2LL * blockIdx.x + t;
      k.red_a[o] = top.a[t], k.red_c[o] = top.c[t], k.red_s[o] = top.s[t], k.red_f[o] = top.d[t];
    }
  }
  __threadfence();
  grid.sync();
  // ---- the 2-rows-per-CTA system, solved by every CTA for itself (the work area is free: no copy is in flight)
  {
    const int m2 = 2 * (int)gridDim.x;
    constexpr int PM = lvl_stride(OL_P1M);
    LevelPtr l1 = sm.w.mid.level(0);
    for (int rr = t; rr < m2; rr += BS_THREADS) {
      const int p = row_slot(rr, m2, PM);
      l1.a[p] = __ldcg(k.red_a + rr), l1.c[p] = __ldcg(k.red_c + rr), l1.s[p] = __ldcg(k.red_s + rr), l1.d[p] = __ldcg(k.red_f + rr);
    }
    sm.w.mid.reduce(t, m2);
    LevelPtr top = sm.w.mid.top();
    if (t == 0) {  // closed 2 x 2 system (a_0 = c_1 = 0), as in mid_body
      const double c0 = top.c[0], s0 = top.s[0], f0 = top.d[0], a1 = top.a[1], s1 = top.s[1], f1 = top.d[1];
      const double det = (s0 * s1 - s0 * a1) - c0 * s1;
      top.d[0] = ((s1 - a1) * f0 - c0 * f1) / det;
      top.d[1] = ((s0 - c0) * f1 - a1 * f0) / det;
    }
    sm.w.mid.backsolve(t, m2);
    if (t < 2) sm.cta.top().d[t] = l1.d[row_slot(2 * (int)blockIdx.x + t, m2, PM)];
  }
  sm.cta.backsolve(t, rowsC);  // starts and ends with a CTA barrier: the work area is free again afterwards
  // ---- pass 2: the tiles again, last one first, swept with the multipliers kept and back-substituted
  if (t == 0 && staged(tb1 - 1)) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the work area was written with ordinary stores
    request(tb1 - 1);
  }
  for (int kk = T - 1; kk >= 0; --kk) {
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
    ChunkGeom gm;
    int rows1;
    bool vec;
    load_tile(tb0 + kk, kk > 0 ? tb0 + kk - 1 : -1, false, A, C, S, F, gm, rows1, vec);
    double u_left = 0.0;
    if (t < bs_rows_left_of(rows1, P1T, OL_NLEV)) u_left = c1.d[row_slot(OL_ROUT * kk + t, rowsC, PC)];
    const typename Pyramid<P1T>::Left left =
        tile_forward<OlSrc<SRC>::FWD, true, OL_NLEV>(sm.w.tile.pyr, k, A, C, S, F, gm.len, gm.g0, t, rows1);
    if (t < left.rows) sm.w.tile.pyr.d[left.pos(t)] = u_left;
    tile_backward<OlSrc<SRC>::BWD, ACCUM, OL_NLEV>(sm.w.tile.pyr, k, A, C, S, F, gm.len, gm.g0, vec, t, rows1);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct OneLaunchPlan {
  int grid = 0;       // CTAs (<= co-resident capacity, <= tiles)
  int p1c = 0;        // capacity class of the CTA level
};

template <int SRC, bool ACCUM, int P1C>
inline int onelaunch_capacity() {
  using Smem = OneLaunchSmem<OlSrc<SRC>::NARR, P1C>;
  int dev = 0, coop = 0, per_sm = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop) return 0;
  if (cudaFuncSetAttribute(k_tri_onelaunch<SRC, ACCUM, P1C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tri_onelaunch<SRC, ACCUM, P1C>, BS_THREADS, sizeof(Smem)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return std::min(per_sm * sms, 8 * OL_P1M / 2);  // the middle system holds 2 rows per CTA in 8 OL_P1M rows
}

// Launches the kernel on (args.k.p0 .. p3, args.k.m); returns 0 on success, -1 when the one-launch kernel does not apply
// (no cooperative launch, system too long for the CTA level), 1 on a CUDA error (bigsys_error()).
// caps: per-device capacities of the instantiations, asked once (index 0: P1C = 72, 1: P1C = 256).
template <int SRC, bool ACCUM>
inline int onelaunch_run(BigSys* s, OneLaunchArgs args, int (&caps)[2], cudaStream_t st, unsigned long long* launches) {
  const size_t m = (size_t)args.k.m;
  const size_t tiles = (m + BS_TILE - 1) / BS_TILE;
  if (tiles < 2) return -1;
  if (caps[0] < 0) caps[0] = onelaunch_capacity<SRC, ACCUM, 72>();
  int which = 0;
  size_t grid = std::min<size_t>(tiles, (size_t)std::max(caps[0], 0));
  if (grid == 0 || (tiles + grid - 1) / grid > 72 / 4) {  // more than 18 tiles per CTA: the big CTA level
    if (caps[1] < 0) caps[1] = onelaunch_capacity<SRC, ACCUM, 256>();
    grid = std::min<size_t>(tiles, (size_t)std::max(caps[1], 0));
    which = 1;
    if (grid == 0 || (tiles + grid - 1) / grid > 256 / 4) return -1;
  }
  CUtensorMap c0, c1, c2, c3;
  int use_tma = s->tma && !(reinterpret_cast<size_t>(args.k.p0) & 15) && bigsys_map(s, args.k.p0, m, &c0);
  if (use_tma && SRC != BS_SRC_NODES)
    use_tma = !((reinterpret_cast<size_t>(args.k.p1) | reinterpret_cast<size_t>(args.k.p2) | reinterpret_cast<size_t>(args.k.p3)) & 15) &&
              bigsys_map(s, args.k.p1, m, &c1) && bigsys_map(s, args.k.p2, m, &c2) && bigsys_map(s, args.k.p3, m, &c3);
  if (!use_tma) std::memset(&c0, 0, sizeof c0);
  if (!use_tma || SRC == BS_SRC_NODES) c1 = c0, c2 = c0, c3 = c0;
  void* params[] = {&c0, &c1, &c2, &c3, &args, &use_tma};
  ++*launches;
  cudaError_t e;
  if (which == 0)
    e = cudaLaunchCooperativeKernel((const void*)k_tri_onelaunch<SRC, ACCUM, 72>, dim3((unsigned)grid), dim3(BS_THREADS), params,
                                    sizeof(OneLaunchSmem<OlSrc<SRC>::NARR, 72>), st);
  else
    e = cudaLaunchCooperativeKernel((const void*)k_tri_onelaunch<SRC, ACCUM, 256>, dim3((unsigned)grid), dim3(BS_THREADS), params,
                                    sizeof(OneLaunchSmem<OlSrc<SRC>::NARR, 256>), st);
  return e == cudaSuccess ? 0 : bigsys_fail(e, "k_tri_onelaunch");
}

}  // namespace dzb
