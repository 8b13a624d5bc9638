// onelaunch_kernels.cuh — assemble + solve of one long bar in ONE launch, the bands never exist in memory.
//
// The reference builds K with its quadrature loop and then runs solve_by_thomas over it (diffusion_solver/
// time_independent.rs:91-196, matrix_solver/mod.rs:17-48).  The multi-launch path of bigsys_kernels.cuh mirrors that
// split: k_assemble_fast writes four bands (32 B/row), the reduce pass reads them, two or three small launches solve the
// reduced system, the solve pass reads the bands again (112 B/row moved for 80 B/row of algorithmic traffic, seven
// dependent launches).  Here the whole thing is one persistent cooperative kernel in which every CTA owns a CONTIGUOUS
// range of 2048-row tiles, so that what its chunks leave behind is a contiguous piece of the next, four times smaller
// system and can be reduced further by the same CTA without waiting for anybody:
//
//   level A (the bar itself, ~34 000 rows per CTA at n = 10^7), per tile: the node coordinates arrive through the TMA (one
//            16 KB box, SWIZZLE_128B, prefetched one tile ahead; the two coordinates either side as bulk copies), every
//            thread forms the 8 rows of its chunk in registers with the lean closed form of assembly_kernels.cuh (ti_elem /
//            ti_row: bit-identical to k_assemble_fast's bands), leaves ONE number per row in memory — the row-sum excess
//            s = lo + di + up, the expensive part of a row — sweeps the chunk and writes its two reduced rows to level B.  No shared-memory pyramid per tile: a pyramid level is ~1 us of single-warp latency during which
//            the rest of the CTA waits (measured: 2 us of every 7 us tile in pass 1, 3.4 of 4.9 us in pass 2);
//   level B (rows / 4) and level C (rows / 16): the same chunk sweep by all 256 threads over the CTA's slice of the
//            previous level's rows (global scratch, L2 resident), no barrier inside a level;
//   level D (rows / 64 <= 576): the CTA's shared-memory pyramid down to its 2 interface rows, ONE grid-wide barrier, then
//            every CTA solves the 2-rows-per-CTA system redundantly in shared memory (592 rows on a B200) — no second
//            barrier, no CTA waits for another one again;
//   and back: D in shared memory, then C, B, A: each chunk is swept again with its multipliers kept in registers and
//            back-substituted from the two values the level above left for it.  Level A reads the coordinates and the
//            excesses through the TMA, tiles in reverse order (what was written last is the likeliest to still be in L2): the
//            off-diagonals are a reciprocal and a multiply-add away from the coordinates (lo = mu K0/h_left + b hp,
//            up = mu K0/h_right + b hn, 1/h by ti_elem's own expression: the same bits as in pass 1).  Pass 1 used to leave 1/h
//            as well; it is write-bandwidth bound once the assembly is taken out (measured: 60 of its 90 us), and the 8 B per
//            row were worth more than the nine reciprocals per chunk they saved pass 2.
//
// HBM traffic: x in twice (16) + s out and in again (16) + u out (8) = 40 B/row, plus the level-B rows and solutions
// (~20 B/row, mostly L2 hits), instead of 112; one launch instead of seven; the rows are formed once.  A mesh with a row the
// closed form does not cover (degenerate spacing: ti_elem's sufficient test fails) is reported through class_flags[1] and
// the caller goes back to materialised bands.
#pragma once
#include "assembly_kernels.cuh"
#include "bigsys_kernels.cuh"
#include "ol_layout.h"

namespace dzb {

constexpr int OL_P1M = 112;                                        // middle system: 2 rows per CTA, up to 896 rows (448 CTAs)
constexpr int OL_TAB_CAP = 256;                                    // rule tables live in shared memory: at most this many nodes
constexpr int OL_ROWS_B = BS_TILE / 4, OL_ROWS_C = BS_TILE / 16, OL_ROWS_D = BS_TILE / 64;  // rows a full tile leaves at each level

struct OneLaunchArgs {
  TileArgs k;        // p0 = x with x[r] the node of row r; m rows; red_* : [2 gridDim.x] scratch; out; flags; bc0 / bc1
  RuleView rule;
  AsmParams P;
  long long n_local; // nodes of the local coordinate array (rows + halo nodes of an open block)
  int halo_left;     // 1 when x[-1] exists (open left end): local node index of row r = r + halo_left
  double* s_ex;      // [m] scratch: row-sum excess of row r
  double* gb[4];     // [OL_ROWS_B tiles] x (a, c, s, f): level B rows
  double* gc[4];     // [OL_ROWS_C tiles] x (a, c, s, f): level C rows
  double *ub, *uc;   // their solutions
  unsigned long long* trace;  // tuning aid (DZB_OL_TRACE=1): [8 gridDim.x] %globaltimer at the phase boundaries of every CTA; nullptr normally
  int* class_flags;  // [0]: k_tri_classify's bits or'ed over every interior row (bit 0: not M-like, bit 1: not mass-like);
                     // [1]: != 0 when some row is outside the closed form.  nullptr: already known for these coordinates
  const int* perm;   // [gridDim.x] position of CTA blockIdx.x in the bar (nullptr: blockIdx.x): the host lays the CTAs out so that
                     // the two of an SM hold NEIGHBOURING tile ranges, one tile longer and one shorter (onelaunch_placement)
  int* probe_smid;   // != nullptr: placement probe — every CTA reports the SM it runs on and returns
  PeerBox px;        // world > 1: the system is one block of a bar split across GPUs; its two interface rows go through the
                     // peer mailboxes (comm_kernels.cuh) between the passes and every rank solves the 2-rows-per-block system
};

// A small tridiagonal system in shared memory, reduced to its first and last row by CYCLIC REDUCTION, in place.
// The pyramids of bigsys_kernels.cuh reduce 8 rows to 2 per level with one thread sweeping a chunk serially: five levels of
// ~1 us each for the ~550 rows a CTA is left with, and again for the 2-rows-per-CTA system — measured 8 + 13.6 us forward and
// ~9 us back per solve, with seven warps of eight parked at barriers.  Here a level halves the system with ALL rows at once:
// at stride st the rows (2q + 1) st are eliminated from their kept neighbours 2q st and (2q + 2) st (one barrier, a dozen
// multiply-adds and two reciprocals per kept row), ~log2(m) levels of ~0.15 us.  Row m-1 is always kept — it carries the
// coupling to whatever follows the system (the next CTA, the next block of a split bar) exactly like row 0 carries the one
// to what precedes it — so the result has the shape the callers already use: two rows coupling (u_0, u_{m-1}) to each other
// and to the outside.  Rows stay in the row-excess form (a, c, s, f), diagonal = s - a - c: for an M-matrix (a, c <= 0 < diagonal)
// the multipliers -a_i / d_j are non-negative and every update adds numbers of one sign, as in the chunk sweeps.
// Eliminated rows keep their entries for the back-substitution, which writes u over f (d[]), last level first.
// Row r lives at slot at(r) = r + r / 16: the strides 2 st of the levels are powers of two, which without the padding would
// put every active row of a level into the same bank (measured: ~0.7 us per level instead of ~0.2).
template <int CAP>
struct CrSystem {
  static constexpr int SLOTS = CAP + CAP / 16 + 1;
  double a[SLOTS], c[SLOTS], s[SLOTS], d[SLOTS];
  __device__ __forceinline__ static int at(int r) { return r + (r >> 4); }
  // the kept neighbour to the right of active row j at stride st
  __device__ __forceinline__ static int right_of(int j, int st, int m) { return j + st < m - 1 ? j + st : m - 1; }
  __device__ __forceinline__ void reduce(int t, int m) {
    for (int lg = 0; (1 << lg) < m - 1; ++lg) {  // stride st = 2^lg (shifts: an integer division per level would cost as much as the level)
      const int st = 1 << lg;
      __syncthreads();
      const int nk = ((m - 2) >> (lg + 1)) + 1;  // regular kept rows 0, 2 st, ... (<= m - 2); slot nk is row m-1
      for (int q = t; q <= nk; q += BS_THREADS) {
        int i, jl = -1, jr = -1;
        if (q < nk) {
          i = q << (lg + 1);
          if (q > 0) jl = i - st;
          if (i + st < m - 1) jr = i + st;
        } else {
          i = m - 1;
          const int mult = (m - 2) >> lg;  // its neighbour to the left at this stride is mult st: eliminated iff mult is odd
          if (mult & 1) jl = mult << lg;
        }
        const int pi = at(i);
        double ai = a[pi], ci = c[pi], si = s[pi], fi = d[pi];
        if (jl >= 0) {
          const int pj = at(jl);
          const double aj = a[pj], cj = c[pj], sj = s[pj];
          const double k1 = -ai * pivot_rcp((sj - aj) - cj);
          si = fma(k1, sj, si), fi = fma(k1, d[pj], fi), ai = k1 * aj;
        }
        if (jr >= 0) {
          const int pj = at(jr);
          const double aj = a[pj], cj = c[pj], sj = s[pj];
          const double k2 = -ci * pivot_rcp((sj - aj) - cj);
          si = fma(k2, sj, si), fi = fma(k2, d[pj], fi), ci = k2 * cj;
        }
        a[pi] = ai, c[pi] = ci, s[pi] = si, d[pi] = fi;  // (nobody reads a kept row at this level but its own thread)
      }
    }
    __syncthreads();
  }
  // d[at(0)] and d[at(m-1)] hold u_0 and u_{m-1}; every other d[at(j)] becomes u_j
  __device__ __forceinline__ void backsolve(int t, int m) {
    int st = 1;
    while (2 * st < m - 1) st <<= 1;  // the last stride reduce() used (none when m <= 2)
    for (; st >= 1 && m > 2; st >>= 1) {
      __syncthreads();
      for (int j = (2 * t + 1) * st; j < m - 1; j += 2 * BS_THREADS * st) {
        const int pj = at(j);
        const double aj = a[pj], cj = c[pj];
        const double num = fma(-cj, d[at(right_of(j, st, m))], fma(-aj, d[at(j - st)], d[pj]));
        d[pj] = num * pivot_rcp((s[pj] - aj) - cj);
      }
    }
    __syncthreads();
  }
};

template <int P1C>
struct OneLaunchSmem {
  union Work {
    alignas(1024) double stage[2][BS_TILE];       // TMA stage.  level A forward: x in [0]; backward: 1/h in [0], excess in [1]
    CrSystem<8 * OL_P1M> mid;                     // only between the two passes, when no copy is in flight
  } w;
  // level D.  With two or more chunk levels (depth >= 2) the system is dead while level A runs — its rows arrive after the
  // forward pass and its solutions have been handed to global scratch before the backward pass — so its space doubles as
  // the backward pass's second stage (stage2()): two tiles in flight per CTA without another 32 KB taken from the L1
  alignas(1024) CrSystem<8 * P1C> cta;
  static_assert(sizeof(CrSystem<8 * P1C>) >= 2 * BS_TILE * sizeof(double), "the second stage lives in the level-D system");
  __device__ __forceinline__ double* stage2(int which) { return reinterpret_cast<double*>(&cta) + which * BS_TILE; }
  double4 uv[OL_TAB_CAP + 1];
  double qx[OL_TAB_CAP];
  unsigned short guessf[GUESS_BINS + 2];
  alignas(16) double halo[8];                     // the two coordinates either side of a staged tile ([4..7]: of the tile in the second stage)
  double xrows[MB_MAX * 8], xends[2 * MB_MAX], xwork[4 * MB_MAX];  // split bar: everybody's interface rows, their solution
  alignas(8) unsigned long long mbar, mbar2;     // mbar2: the second stage of the backward pass
};
// 16 or 32 bytes next to a staged tile, through the same mbarrier (plain bulk copy: 16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<unsigned long long>(src)), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// element e (0 .. 2047) of a staged tile: 128-byte tensor rows, 16-byte columns XOR-swizzled with the row (SWIZZLE_128B)
__device__ __forceinline__ double stage_elem(const double* stage, int e) {
  const int row = e >> 4, col = (e >> 1) & 7;
  return *reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(stage) + row * 128 + ((col ^ (row & 7)) << 4) + (e & 1) * 8);
}
// the 8 consecutive elements 8 t .. 8 t + 7 of a staged tile (chunk t = bytes [64 t, 64 t + 64) = 16-byte columns 4 (t & 1) + j
// of tensor row t / 2): four conflict-free LDS.128
__device__ __forceinline__ void stage_chunk(const double* stage, int t, double* v) {
  const int row = t >> 1;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(stage) + row * 128;
#pragma unroll
  for (int j = 0; j < BS_R / 2; ++j) {
    const int col = ((((t & 1) << 2) + j) ^ (row & 7)) << 4;
    const double2 q = *reinterpret_cast<const double2*>(base + col);
    v[2 * j] = q.x, v[2 * j + 1] = q.y;
  }
}

__device__ __forceinline__ void ol_stamp(const OneLaunchArgs& a, int slot) {
  if (a.trace && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    if (slot == 0) {  // the low 12 bits of the start stamp carry the SM the CTA runs on
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      ns = (ns & ~0xfffULL) | (smid & 0xfffu);
    }
    a.trace[8 * blockIdx.x + slot] = ns;
  }
}

struct OlTables {  // rule tables, in shared memory (the launcher refuses rules with more than OL_TAB_CAP nodes)
  const double4* __restrict__ uv;
  const unsigned short* __restrict__ guessf;
  const double* __restrict__ qx;
  int m;
};

// Rows i0 .. i0+7 (local node indices) of the static diffusion system in row-excess form (a, c, s, f), from the
// coordinates X[k] = x_local[i0 - 2 + k] (entries outside [0, n_local) are zero and never reach a result).
// Straight-line code: rows that do not exist, Dirichlet rows and rows outside the closed
// form are sorted out with selects, so the eight rows of a chunk interleave freely.
template <bool CLASSIFY>
__device__ __forceinline__ void ti_rows_from_nodes(const double (&X)[12], long long i0, int len, const OneLaunchArgs& a, const TiK& K,
                                                   const OlTables& tb, double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R],
                                                   double (&F)[BS_R], int& cls, bool& bad) {
  TiElem E[BS_R + 1];  // element el between X[el + 1] and X[el + 2]
#pragma unroll
  for (int el = 0; el <= BS_R; ++el) E[el] = ti_elem(X[el + 1], X[el + 2], K);
  // j-ranges instead of 64-bit comparisons per row: rows [jd0] / [jd1] are the Dirichlet rows when they fall into this chunk,
  // rows jlo <= j < jhi are interior rows, the outer-neighbour tests apply from jl2 on / below jh2
  const long long nl = a.n_local;
  const int jd0 = (int)max(-1LL, min(8LL, -i0)), jd1 = (int)max(-1LL, min(8LL, nl - 1 - i0));
  const int jlo = max(0, jd0 + 1), jhi = min(len, jd1);
  const int jl2 = (int)max(0LL, min(9LL, 2 - i0)), jh2 = (int)max(-1LL, min(8LL, nl - 2 - i0));  // i >= 2 / i + 2 < nl
#pragma unroll
  for (int j = 0; j < BS_R; ++j) {
    const bool inter = j >= jlo && j < jhi;
    // whether the closed form covers the row is a property of the coordinates: looked at (together with the matrix class) the
    // first time these coordinates are solved for, taken for granted afterwards
    bool ok = true;
    if (CLASSIFY) {
      ok = E[j].ok && E[j + 1].ok;
      ok = ok && (j < jl2 || !(X[j + 1] < X[j]));
      ok = ok && (j >= jh2 || !(X[j + 4] < X[j + 3]));
      bad = bad || (inter && !ok);
    }
    double lo, di, up;
    ti_row(E[j], E[j + 1], X[j + 1], X[j + 2], X[j + 3], K, tb.uv, tb.guessf, tb.qx, tb.m, lo, di, up, inter && ok);
    const double ex = excess3(lo, di, up);
    if (CLASSIFY && inter) {
      if (!(lo <= 0.0 && up <= 0.0 && ex >= -1e-6 * di)) cls |= 1;
      if (!(lo >= 0.0 && up >= 0.0 && di >= 1.25 * (lo + up))) cls |= 2;
    }
    A[j] = lo, C[j] = up, S[j] = ex, F[j] = 0.0;
  }
  if (jlo != 0 || jhi != BS_R) {  // a chunk with padding or Dirichlet rows, (0, 1, 0, f) with excess 1: the first and the last of the bar
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const bool inter = j >= jlo && j < jhi;
      A[j] = inter ? A[j] : 0.0, C[j] = inter ? C[j] : 0.0, S[j] = inter ? S[j] : 1.0;
      F[j] = (j == jd0 && j < len) ? a.P.bc0 : ((j == jd1 && j < len) ? a.P.bc1 : 0.0);
    }
  }
}

// The chunk of thread t: rows g0 .. g0+len of the system.  Staged tiles hand over the 8 own coordinates from shared
// memory and the 2 + 2 halo coordinates from the neighbouring lanes; everything else reads global memory directly.
__device__ __forceinline__ void nodes_from_stage(const double* stage, const double* halo, bool has_halo, const OneLaunchArgs& a,
                                                 long long g0, int t, double (&X)[12]) {
  stage_chunk(stage, t, &X[2]);
  const int lane = t & 31;
  X[0] = __shfl_up_sync(0xffffffffu, X[8], 1);
  X[1] = __shfl_up_sync(0xffffffffu, X[9], 1);
  X[10] = __shfl_down_sync(0xffffffffu, X[2], 1);
  X[11] = __shfl_down_sync(0xffffffffu, X[3], 1);
  const long long i0 = g0 + a.halo_left;  // local node index of the chunk's first row
  const double* xl = a.k.p0 - a.halo_left;
  if (lane == 0) {
    if (t > 0) X[0] = stage_elem(stage, 8 * t - 2), X[1] = stage_elem(stage, 8 * t - 1);
    else if (has_halo) X[0] = halo[0], X[1] = halo[1];
    else X[0] = i0 >= 2 ? xl[i0 - 2] : 0.0, X[1] = i0 >= 1 ? xl[i0 - 1] : 0.0;
  }
  if (lane == 31) {
    if (t < BS_THREADS - 1) X[10] = stage_elem(stage, 8 * t + 8), X[11] = stage_elem(stage, 8 * t + 9);
    else if (has_halo) X[10] = halo[2], X[11] = halo[3];
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

// Pass 2: rows g0 .. g0+len again, from the coordinates X[k] = x_local[i0 - 2 + k] and the excesses pass 1 left behind.
__device__ __forceinline__ void ti_rows_from_scratch(const double (&X)[12], const double (&Sv)[BS_R], long long i0, int len,
                                                     const OneLaunchArgs& a, const TiK& K, double (&A)[BS_R], double (&C)[BS_R],
                                                     double (&S)[BS_R], double (&F)[BS_R]) {
  double R[BS_R + 1];  // 1/h of the element between X[el + 1] and X[el + 2]: ti_elem's expression, the bits of pass 1
#pragma unroll
  for (int el = 0; el <= BS_R; ++el) R[el] = rcp_newton(X[el + 2] - X[el + 1]);
  const long long nl = a.n_local;
  const int jd0 = (int)max(-1LL, min(8LL, -i0)), jd1 = (int)max(-1LL, min(8LL, nl - 1 - i0));
  const int jlo = max(0, jd0 + 1), jhi = min(len, jd1);
#pragma unroll
  for (int j = 0; j < BS_R; ++j) {
    const double lo = fma(K.mu, K.K0 * R[j], K.bhp), up = fma(K.mu, K.K0 * R[j + 1], K.bhn);  // ti_row's expressions
    A[j] = lo, C[j] = up, S[j] = Sv[j], F[j] = 0.0;
  }
  if (jlo != 0 || jhi != BS_R) {  // as in ti_rows_from_nodes
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const bool inter = j >= jlo && j < jhi;
      A[j] = inter ? A[j] : 0.0, C[j] = inter ? C[j] : 0.0, S[j] = inter ? S[j] : 1.0;
      F[j] = (j == jd0 && j < len) ? a.P.bc0 : ((j == jd1 && j < len) ? a.P.bc1 : 0.0);
    }
  }
}


// ---- levels B and C: rows (a, c, s, f) in global scratch ---------------------------------------------------------------
// rows r0 .. r0+len of the arrays g[0..3] -> registers (L2 loads: the rows were written by this CTA moments ago)
// with_f == false: the chunk's right-hand sides are known to be zero (see f_chunk) and are not read.
__device__ __forceinline__ void rows_load(double* const (&g)[4], long long r0, int len, double (&A)[BS_R], double (&C)[BS_R],
                                          double (&S)[BS_R], double (&F)[BS_R], bool with_f) {
  if (len == BS_R && !(r0 & 1)) {
    const double2* q0 = reinterpret_cast<const double2*>(g[0] + r0);
    const double2* q1 = reinterpret_cast<const double2*>(g[1] + r0);
    const double2* q2 = reinterpret_cast<const double2*>(g[2] + r0);
    const double2* q3 = reinterpret_cast<const double2*>(g[3] + r0);
#pragma unroll
    for (int j = 0; j < BS_R / 2; ++j) {
      const double2 v0 = __ldca(q0 + j), v1 = __ldca(q1 + j), v2 = __ldca(q2 + j);
      const double2 v3 = with_f ? __ldca(q3 + j) : make_double2(0.0, 0.0);
      A[2 * j] = v0.x, A[2 * j + 1] = v0.y, C[2 * j] = v1.x, C[2 * j + 1] = v1.y;
      S[2 * j] = v2.x, S[2 * j + 1] = v2.y, F[2 * j] = v3.x, F[2 * j + 1] = v3.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < BS_R; ++j) {
      const bool in = j < len;
      A[j] = in ? __ldca(g[0] + r0 + j) : 0.0, C[j] = in ? __ldca(g[1] + r0 + j) : 0.0;
      S[j] = in ? __ldca(g[2] + r0 + j) : 1.0, F[j] = (in && with_f) ? __ldca(g[3] + r0 + j) : 0.0;
    }
  }
}
// The levels B and C are private to the CTA (written and read by its own threads, never by another SM), so their loads may go
// through the L1 — and a thread can ask for the lines of its NEXT chunk before it sweeps this one: the sweeps of these levels
// were waits for the L2 (ncu: long_sb + lg on their first use of a loaded row, ~9 % of the kernel's stall samples).
__device__ __forceinline__ void rows_prefetch(double* const (&g)[4], long long r0, bool with_f) {
#pragma unroll
  for (int k = 0; k < 3; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(g[k] + r0));
  if (with_f) asm volatile("prefetch.global.L1 [%0];" ::"l"(g[3] + r0));
}
// the two reduced rows of a chunk -> rows q, q + 1 (q even) of the arrays g[0..3]
__device__ __forceinline__ void rows_store2(double* const (&g)[4], long long q, const double (&o)[8], bool with_f) {
#pragma unroll
  for (int k = 0; k < 3; ++k) *reinterpret_cast<double2*>(g[k] + q) = make_double2(o[k], o[4 + k]);
  if (with_f) *reinterpret_cast<double2*>(g[3] + q) = make_double2(o[3], o[7]);
}
// The right-hand side of the static problem is zero but for the two Dirichlet rows (time_independent.rs:168-182), and a chunk
// of zero right-hand sides reduces to zero right-hand sides exactly.  At every level only the first chunk of the CTA that holds
// the bar's first row and the last chunk of the CTA that holds its last row carry a right-hand side: nobody else writes or reads
// the f arrays of the levels B and C (a quarter of their traffic; the launcher's scratch is zero-filled once, and positions
// nobody writes stay zero).  An open end (a block of a split bar) has no Dirichlet row.
struct FChunks {
  bool first, last;  // this CTA holds the closed first / last row of the system
  __device__ __forceinline__ bool operator()(int c, int nc) const { return (first && c == 0) || (last && c == nc - 1); }
};
template <bool KEEP>
__device__ __forceinline__ void chunk_sweep(double (&A)[BS_R], double (&C)[BS_R], double (&S)[BS_R], double (&F)[BS_R], int len,
                                            double (&o)[8]) {
  if (len == BS_R) chunk0_sweep<KEEP, true>(A, C, S, F, len, o);
  else chunk0_sweep<KEEP, false>(A, C, S, F, len, o);
}
// One streaming level, forward: the CTA's m rows at src[0..3][base ..] -> 2 rows per chunk at dst[0..3][dbase ..]
__device__ __forceinline__ void level_forward(double* const (&src)[4], long long base, int m, double* const (&dst)[4], long long dbase, int t,
                                              const FChunks& fc) {
  const int nc = (m + BS_R - 1) / BS_R;
  for (int c = t; c < nc; c += BS_THREADS) {
    const ChunkGeom gm = chunk_geom(m, c, 0);
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R], o[8];
    if (c + BS_THREADS < nc) rows_prefetch(src, base + chunk_geom(m, c + BS_THREADS, 0).g0, fc(c + BS_THREADS, nc));
    rows_load(src, base + gm.g0, gm.len, A, C, S, F, fc(c, nc));
    chunk_sweep<false>(A, C, S, F, gm.len, o);
    rows_store2(dst, dbase + 2 * c, o, fc(c, nc));  // (its image is chunk c / 4 of the next level: first / last again)
  }
}
// ... and backward: usrc[ubase + 2 c], [.. + 1] hold u of the chunk's first / last row; u of all its rows -> udst[base ..]
__device__ __forceinline__ void level_backward(double* const (&src)[4], long long base, int m, const double* usrc, long long ubase,
                                               double* udst, int t, const FChunks& fc) {
  const int nc = (m + BS_R - 1) / BS_R;
  for (int c = t; c < nc; c += BS_THREADS) {
    const ChunkGeom gm = chunk_geom(m, c, 0);
    const double2 ends = __ldcg(reinterpret_cast<const double2*>(usrc + ubase + 2 * c));
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R], o[8], u[BS_R];
    if (c + BS_THREADS < nc) rows_prefetch(src, base + chunk_geom(m, c + BS_THREADS, 0).g0, fc(c + BS_THREADS, nc));
    rows_load(src, base + gm.g0, gm.len, A, C, S, F, fc(c, nc));
    chunk_sweep<true>(A, C, S, F, gm.len, o);
    chunk0_backward(A, C, S, F, gm.len, ends.x, ends.y, u);
    double* dst = udst + base + gm.g0;
    if (gm.len == BS_R && !((base + gm.g0) & 1)) {
#pragma unroll
      for (int j = 0; j < BS_R / 2; ++j) reinterpret_cast<double2*>(dst)[j] = make_double2(u[2 * j], u[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < BS_R; ++j)
        if (j < gm.len) dst[j] = u[j];
    }
  }
}

template <int P1C, int MINB>
__global__ void __launch_bounds__(BS_THREADS, MINB) k_tri_onelaunch(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_s, const OneLaunchArgs a,
                                                                 const int use_tma) {
  using Smem = OneLaunchSmem<P1C>;
  extern __shared__ __align__(1024) unsigned char ol_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(ol_smem_raw);
  if (smem_u32(ol_smem_raw) & 1023) __trap();
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const TileArgs& k = a.k;
  const int t = threadIdx.x;
  if (a.probe_smid) {
    if (t == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      a.probe_smid[blockIdx.x] = (int)smid;
    }
    return;
  }
  const int vb = a.perm ? a.perm[blockIdx.x] : (int)blockIdx.x;  // the CTA's position along the bar
  const long long ntiles = (k.m + BS_TILE - 1) / BS_TILE;
  // Contiguous tile ranges, an even share each.  (The two CTAs of an SM do not finish pass 1 together: the one placed first wins
  // the issue slots and is done after ~75 us at 10^7 rows, its neighbour after ~95 us.  Giving the first arrival — found with a
  // per-%smid arrival counter, blockIdx does not tell — a larger share evened the two out and changed nothing: the SM takes as
  // long for its 33 tiles either way, so the split stays even.)  What does matter is how many tiles an SM gets IN TOTAL: the
  // block scheduler of a B200 puts CTAs b and b + 142 (b + 6 on six SMs) on one SM — the same in every launch on every box
  // looked at — and with 16.5 tiles per CTA laid out by blockIdx both of them got 16 or both 17: half the SMs had 34 tiles, half
  // 32, and everybody waited for the 34s, at the grid barrier and at the end.  The host therefore probes the placement once
  // (probe_smid) and hands each CTA a position `vb` such that the two CTAs of an SM are neighbours along the bar: 33 tiles per SM.  A ragged last tile takes the direct-load path, which costs about two staged
  // tiles, and everybody would wait for its CTA at the grid barrier: it counts twice in the split
  long long tb0, tb1;
  {
    const long long units = ntiles + ((use_tma && (k.m % BS_TILE)) ? 1 : 0);
    tb0 = min(ntiles, (long long)vb * units / gridDim.x);
    tb1 = vb + 1 == (int)gridDim.x ? ntiles : min(ntiles, ((long long)vb + 1) * units / gridDim.x);
  }
  const int T = (int)(tb1 - tb0);  // >= 1: the launcher never starts more CTAs than tiles
  const FChunks fc{tb0 == 0 && !(k.flags & BS_OPEN_LEFT), tb1 == ntiles && !(k.flags & BS_OPEN_RIGHT)};
  // the CTA's systems: level A rows [R0, R0 + mA) of the bar, then mB = 2 chunks(mA) rows at gb[..][oB ..], mC, mD (shared memory)
  const long long R0 = tb0 * BS_TILE;
  const int mA = (int)(min(k.m, tb1 * BS_TILE) - R0);
  const int nA = (mA + BS_R - 1) / BS_R, mB = 2 * nA, nB = (mB + BS_R - 1) / BS_R, mC = 2 * nB, nC = (mC + BS_R - 1) / BS_R, mD = 2 * nC;
  const long long oB = tb0 * OL_ROWS_B, oC = tb0 * OL_ROWS_C;
  // How many chunk levels the CTA goes through before its system fits the shared-memory cyclic reduction (8 P1C rows):
  // 1: level A's reduced rows go there directly (no global scratch at all: up to 8 P1C / 512 tiles per CTA);
  // 2: level B is swept from global scratch into it; 3: levels B and C.  A CTA-local choice.
  const int depth = mB <= 8 * P1C ? 1 : (mC <= 8 * P1C ? 2 : 3);
  const int mS = depth == 1 ? mB : (depth == 2 ? mC : mD);  // rows of the CTA's shared-memory system
  const TiK K = make_tik(a.rule, a.P);
  const OlTables tb{sm.uv, sm.guessf, sm.qx, a.rule.m};
  const bool classify = a.class_flags != nullptr;
  auto staged = [&](long long tile) { return use_tma && (tile + 1) * BS_TILE <= k.m; };  // (m % BS_TILE == 1 is refused by the launcher)
  unsigned phase = 0;

  // ---- level A forward: assemble, leave (1/h, s) per row, two reduced rows per chunk -> level B
  // the two coordinates either side of a staged tile ride along as two 16-byte bulk copies when both pairs exist (every tile
  // but the first and, at most, the last full one): nobody then waits for a global load on a tile's critical path
  auto x_halo = [&](long long tile) { return tile > 0 && (tile + 1) * BS_TILE + 1 + a.halo_left < a.n_local; };
  auto request_x = [&](long long tile) {  // thread 0
    const bool hh = x_halo(tile);
    // The stage was just READ with ordinary loads by every thread (ordered before this point by the CTA barrier) and is about
    // to be WRITTEN by the async proxy: without a proxy fence from the issuing thread a tile now and then kept some 128-byte
    // lines of its predecessor (measured on a B200: one CTA in ~300 per solve formed rows from coordinates 2048 nodes off; on
    // a regular bar, where K 1 = 0, that is a wrong answer, on a jittered one a 1e-6 blemish).
    asm volatile("fence.proxy.async;" ::: "memory");
    mbar_arrive_expect_tx(&sm.mbar, (unsigned)(BS_TILE * sizeof(double)) + (hh ? 32u : 0u));
    tma_load_box(sm.w.stage[0], &map_x, (int)(tile * BS_TMA_BOX), &sm.mbar);
    if (hh) {
      bulk_load(sm.halo, k.p0 + tile * BS_TILE - 2, 16, &sm.mbar);
      bulk_load(sm.halo + 2, k.p0 + (tile + 1) * BS_TILE, 16, &sm.mbar);
    }
  };
  // the first tile is on its way while the CTA copies the rule tables (the barrier after them also publishes the mbarriers)
  if (t == 0) {
    mbar_init(&sm.mbar, 1), mbar_init(&sm.mbar2, 1);
    if (staged(tb0)) request_x(tb0);
  }
  for (int j = t; j <= a.rule.m; j += BS_THREADS) sm.uv[j] = a.rule.uv[j];
  for (int j = t; j < a.rule.m; j += BS_THREADS) sm.qx[j] = a.rule.x[j];
  for (int j = t; j < GUESS_BINS + 2; j += BS_THREADS) sm.guessf[j] = a.rule.guessf[j];
  __syncthreads();
  ol_stamp(a, 0);
  int cls = 0;
  bool bad = false;
  for (int kk = 0; kk < T; ++kk) {
    const long long tile = tb0 + kk;
    const int c = kk * BS_THREADS + t;
    const ChunkGeom gm = chunk_geom(mA, c, 0);  // len == 0: no such chunk
    const long long G0 = R0 + gm.g0;            // first row of the chunk
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
    {
      double X[12];
      if (staged(tile)) {
        mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        nodes_from_stage(sm.w.stage[0], sm.halo, x_halo(tile), a, G0, t, X);
      } else {
        nodes_direct(a, G0, X);
      }
      __syncthreads();  // every chunk is in registers: the stage can take the next tile
      if (t == 0 && kk + 1 < T && staged(tile + 1)) request_x(tile + 1);
      if (classify) ti_rows_from_nodes<true>(X, G0 + a.halo_left, gm.len, a, K, tb, A, C, S, F, cls, bad);
      else ti_rows_from_nodes<false>(X, G0 + a.halo_left, gm.len, a, K, tb, A, C, S, F, cls, bad);
      if (gm.len == BS_R && !(G0 & 1)) {
        double2* ps = reinterpret_cast<double2*>(a.s_ex + G0);
#pragma unroll
        for (int j = 0; j < BS_R / 2; ++j) ps[j] = make_double2(S[2 * j], S[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < BS_R; ++j)
          if (j < gm.len) a.s_ex[G0 + j] = S[j];
      }
    }
    if (gm.len >= 2) {
      double o[8];
      chunk_sweep<false>(A, C, S, F, gm.len, o);
      if (depth == 1) {
        const int p0 = sm.cta.at(2 * c), p1 = sm.cta.at(2 * c + 1);
        sm.cta.a[p0] = o[0], sm.cta.c[p0] = o[1], sm.cta.s[p0] = o[2], sm.cta.d[p0] = o[3];
        sm.cta.a[p1] = o[4], sm.cta.c[p1] = o[5], sm.cta.s[p1] = o[6], sm.cta.d[p1] = o[7];
      } else {
        rows_store2(a.gb, oB + 2 * c, o, fc(c, nA));
      }
    }
  }
  if (classify) {
    const int c0 = __syncthreads_or(cls & 1), c1b = __syncthreads_or(cls & 2), cb = __syncthreads_or(bad);
    if (t == 0) {
      if (c0 | c1b) atomicOr(a.class_flags, (c0 ? 1 : 0) | (c1b ? 2 : 0));
      if (cb) atomicOr(a.class_flags + 1, 1);
    }
  }
  ol_stamp(a, 1);
  // ---- levels B and C forward as far as needed (no barrier inside a level), then the CTA's system in shared memory down to
  // its two interface rows
  __syncthreads();
  if (depth == 3) {
    level_forward(a.gb, oB, mB, a.gc, oC, t, fc);
    __syncthreads();
  }
  ol_stamp(a, 6);
  if (depth >= 2) {  // the last global level -> shared memory
    double* const(&src)[4] = depth == 3 ? a.gc : a.gb;
    const long long base = depth == 3 ? oC : oB;
    const int mL = depth == 3 ? mC : mB, nL = depth == 3 ? nC : nB;
    for (int c = t; c < nL; c += BS_THREADS) {
      const ChunkGeom gm = chunk_geom(mL, c, 0);
      double A[BS_R], C[BS_R], S[BS_R], F[BS_R], o[8];
      if (c + BS_THREADS < nL) rows_prefetch(src, base + chunk_geom(mL, c + BS_THREADS, 0).g0, fc(c + BS_THREADS, nL));
      rows_load(src, base + gm.g0, gm.len, A, C, S, F, fc(c, nL));
      chunk_sweep<false>(A, C, S, F, gm.len, o);
      const int p0 = sm.cta.at(2 * c), p1 = sm.cta.at(2 * c + 1);
      sm.cta.a[p0] = o[0], sm.cta.c[p0] = o[1], sm.cta.s[p0] = o[2], sm.cta.d[p0] = o[3];
      sm.cta.a[p1] = o[4], sm.cta.c[p1] = o[5], sm.cta.s[p1] = o[6], sm.cta.d[p1] = o[7];
    }
  }
  ol_stamp(a, 7);
  sm.cta.reduce(t, mS);
  if (t < 2) {  // the CTA's two interface rows: its first and its last
    const long long o = 2LL * vb + t;
    const int r = sm.cta.at(t ? mS - 1 : 0);
    k.red_a[o] = sm.cta.a[r], k.red_c[o] = sm.cta.c[r], k.red_s[o] = sm.cta.s[r], k.red_f[o] = sm.cta.d[r];
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");  // writers' side of the cross-proxy ordering (see request_rs): both are needed
  __threadfence();
  ol_stamp(a, 2);
  grid.sync();
  ol_stamp(a, 3);
  // ---- the 2-rows-per-CTA system, solved by every CTA for itself (the work area is free: no copy is in flight)
  {
    const int m2 = 2 * (int)gridDim.x;
    auto& mid = sm.w.mid;
    for (int rr = t; rr < m2; rr += BS_THREADS) {
      const int pr = mid.at(rr);
      mid.a[pr] = __ldcg(k.red_a + rr), mid.c[pr] = __ldcg(k.red_c + rr), mid.s[pr] = __ldcg(k.red_s + rr), mid.d[pr] = __ldcg(k.red_f + rr);
    }
    mid.reduce(t, m2);
    const int last = mid.at(m2 - 1);
    if (a.px.world > 1) {
      // One block of a split bar: its two rows keep their couplings to the neighbouring blocks.  CTA 0 posts them to every
      // rank (every CTA holds the same bits), every CTA collects all of them from the local mailbox and solves the
      // 2-rows-per-block interface system for itself: no launch, no host, one NVLink store + flag per peer.
      if (t < a.px.world) {
        if (blockIdx.x == 0) {
          const double v[8] = {mid.a[0], mid.c[0], mid.s[0], mid.d[0], mid.a[last], mid.c[last], mid.s[last], mid.d[last]};
          mb_post(a.px, t, v);
        }
        mb_wait(a.px, t, sm.xrows + 8 * t);
      }
      __syncthreads();
      if (t == 0) {
        interface_solve_serial(sm.xrows, 2 * a.px.world, sm.xends, sm.xwork);
        mid.d[0] = sm.xends[2 * a.px.rank], mid.d[last] = sm.xends[2 * a.px.rank + 1];
      }
    } else if (t == 0) {  // closed 2 x 2 system (a_0 = c_last = 0), as in mid_body
      const double c0 = mid.c[0], s0 = mid.s[0], f0 = mid.d[0], a1 = mid.a[last], s1 = mid.s[last], f1 = mid.d[last];
      const double det = (s0 * s1 - s0 * a1) - c0 * s1;
      mid.d[0] = ((s1 - a1) * f0 - c0 * f1) / det;
      mid.d[last] = ((s0 - c0) * f1 - a1 * f0) / det;
    }
    mid.backsolve(t, m2);
    if (t < 2) sm.cta.d[sm.cta.at(t ? mS - 1 : 0)] = mid.d[mid.at(2 * vb + t)];
  }
  sm.cta.backsolve(t, mS);  // starts and ends with a CTA barrier: the work area is free again afterwards
  ol_stamp(a, 4);
  const bool two_stages = depth >= 2;
  auto request_rs = [&](long long tile, int b) {  // thread 0
    // as in request_x; here the source, too, was written with ordinary stores (the forward pass, other threads, another CTA
    // for the value left of the CTA's first tile): generic -> async ordering for both state spaces, from the issuing thread,
    // after it has synchronised with the writers (the grid barrier and the CTA barriers since)
    asm volatile("fence.proxy.async;" ::: "memory");
    const bool hh = x_halo(tile);
    unsigned long long* mb = b ? &sm.mbar2 : &sm.mbar;
    mbar_arrive_expect_tx(mb, (unsigned)(2 * BS_TILE * sizeof(double)) + (hh ? 32u : 0u));
    const int row = (int)(tile * BS_TMA_BOX);
    tma_load_box(b ? sm.stage2(0) : sm.w.stage[0], &map_x, row, mb);
    tma_load_box(b ? sm.stage2(1) : sm.w.stage[1], &map_s, row, mb);
    if (hh) {
      bulk_load(sm.halo + 4 * b, k.p0 + tile * BS_TILE - 2, 16, mb);
      bulk_load(sm.halo + 4 * b + 2, k.p0 + (tile + 1) * BS_TILE, 16, mb);
    }
  };
  // The first two tiles of the backward pass are asked for as soon as their stages are free — the work area now (the middle
  // system is solved, and the proxy fence of the request covers its ordinary stores), the level-D system's space once its
  // solutions have been handed down — so that they land while the global levels are swept backward (ncu: the wait for the
  // first tile was 4 % of the kernel's stall samples when the requests followed those sweeps).
  if (t == 0 && staged(tb1 - 1)) request_rs(tb1 - 1, 0);
  // ---- the global levels backward
  if (depth >= 2) {  // shared memory -> the last global level
    double* const(&src)[4] = depth == 3 ? a.gc : a.gb;
    const long long base = depth == 3 ? oC : oB;
    const int mL = depth == 3 ? mC : mB, nL = depth == 3 ? nC : nB;
    double* udst = depth == 3 ? a.uc : a.ub;
    for (int c = t; c < nL; c += BS_THREADS) {
      const ChunkGeom gm = chunk_geom(mL, c, 0);
      const double us = sm.cta.d[sm.cta.at(2 * c)], ue = sm.cta.d[sm.cta.at(2 * c + 1)];
      double A[BS_R], C[BS_R], S[BS_R], F[BS_R], o[8], u[BS_R];
      if (c + BS_THREADS < nL) rows_prefetch(src, base + chunk_geom(mL, c + BS_THREADS, 0).g0, fc(c + BS_THREADS, nL));
      rows_load(src, base + gm.g0, gm.len, A, C, S, F, fc(c, nL));
      chunk_sweep<true>(A, C, S, F, gm.len, o);
      chunk0_backward(A, C, S, F, gm.len, us, ue, u);
      double* dst = udst + base + gm.g0;
      if (gm.len == BS_R && !((base + gm.g0) & 1)) {
#pragma unroll
        for (int j = 0; j < BS_R / 2; ++j) reinterpret_cast<double2*>(dst)[j] = make_double2(u[2 * j], u[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < BS_R; ++j)
          if (j < gm.len) dst[j] = u[j];
      }
    }
    __syncthreads();
    if (t == 0 && T >= 2 && staged(tb1 - 2)) request_rs(tb1 - 2, 1);  // (two_stages: depth >= 2)
  }
  if (depth == 3) {
    level_backward(a.gb, oB, mB, a.uc, oC, a.ub, t, fc);
    __syncthreads();
  }

  // ---- level A backward: the tiles again, last one first.  This pass waits for its tiles if only one is in flight (ncu: a tenth
  // of the kernel's stall samples sat on the mbarrier here — 32 KB per tile against 16 KB forward, and less arithmetic to hide them
  // behind), so with depth >= 2 the tiles alternate between the work area and the dead level-D system: the request for the tile
  // after next goes out as soon as this tile's chunks are in registers
  unsigned phase2 = 0;
  const int ahead = two_stages ? 2 : 1;
  for (int kk = T - 1; kk >= 0; --kk) {
    const long long tile = tb0 + kk;
    const int c = kk * BS_THREADS + t;
    const int b = two_stages ? ((T - 1 - kk) & 1) : 0;  // the stage this tile went to
    const ChunkGeom gm = chunk_geom(mA, c, 0);
    const long long G0 = R0 + gm.g0;
    double2 ends = make_double2(0.0, 0.0);
    if (gm.len >= 2) {
      if (depth == 1) ends = make_double2(sm.cta.d[sm.cta.at(2 * c)], sm.cta.d[sm.cta.at(2 * c + 1)]);
      else ends = __ldcg(reinterpret_cast<const double2*>(a.ub + oB + 2 * c));
    }
    double A[BS_R], C[BS_R], S[BS_R], F[BS_R];
    {
      double X[12], Sv[BS_R];
      if (staged(tile)) {
        if (b) {
          mbar_wait(&sm.mbar2, phase2);
          phase2 ^= 1;
        } else {
          mbar_wait(&sm.mbar, phase);
          phase ^= 1;
        }
        nodes_from_stage(b ? sm.stage2(0) : sm.w.stage[0], sm.halo + 4 * b, x_halo(tile), a, G0, t, X);
        stage_chunk(b ? sm.stage2(1) : sm.w.stage[1], t, Sv);
      } else {
        nodes_direct(a, G0, X);
#pragma unroll
        for (int j = 0; j < BS_R; ++j) Sv[j] = j < gm.len ? __ldcg(a.s_ex + G0 + j) : 1.0;
      }
      __syncthreads();  // every chunk is in registers: the stage can take the tile after next (the next one with a single stage)
      if (t == 0 && kk >= ahead && staged(tile - ahead)) request_rs(tile - ahead, b);
      ti_rows_from_scratch(X, Sv, G0 + a.halo_left, gm.len, a, K, A, C, S, F);
    }
    if (gm.len >= 2) {
      double o[8], u[BS_R];
      chunk_sweep<true>(A, C, S, F, gm.len, o);
      chunk0_backward(A, C, S, F, gm.len, ends.x, ends.y, u);
      // a Dirichlet row is returned exactly as Thomas returns it, f / 1 (matrix_solver/mod.rs:36-39 with zero couplings)
      if (G0 == 0 && !(k.flags & BS_OPEN_LEFT)) u[0] = k.bc0 / 1.0;
      if (G0 + gm.len == k.m && !(k.flags & BS_OPEN_RIGHT)) {
#pragma unroll
        for (int j = 0; j < BS_R; ++j)
          if (j == gm.len - 1) u[j] = k.bc1 / 1.0;
      }
      double* dst = k.out + G0;
      if (gm.len == BS_R && !(reinterpret_cast<size_t>(dst) & 15)) {
#pragma unroll
        for (int j = 0; j < BS_R / 2; ++j) reinterpret_cast<double2*>(dst)[j] = make_double2(u[2 * j], u[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < BS_R; ++j)
          if (j < gm.len) dst[j] = u[j];
      }
    }
  }
  ol_stamp(a, 5);
}

// ---- host side -------------------------------------------------------------------------------------------------------
// Two builds: a shared-memory system of up to 1024 rows (32 tiles per CTA: 1.9 10^7 rows on a B200; up to 2 tiles per CTA without
// any global level, up to 8 with one) at two CTAs per SM (128 registers), of up to 2048 rows (64 tiles) at one CTA per SM for
// devices where two do not fit.  (1536 rows would take the blocks of a bar split over 8 GPUs without a global level, but the
// 17 KB of shared memory come out of the L1: 2-4 % slower at 10^7 rows and beyond.)  (The cyclic reduction needs 34 bytes of shared memory
// per level-D row; with the pyramids the two-per-SM build stopped at 18 tiles per CTA and anything longer ran at one CTA per
// SM, 0.27 instead of ~0.21 ms at 1.2 10^7 rows.)
// (Three CTAs per SM with 80 registers spill and are slower: 0.275 against 0.209 ms at 10^7 rows.)
template <int P1C, int MINB>
inline int onelaunch_capacity() {
  using Smem = OneLaunchSmem<P1C>;
  int dev = 0, coop = 0, per_sm = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop) return 0;
  if (cudaFuncSetAttribute(k_tri_onelaunch<P1C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tri_onelaunch<P1C, MINB>, BS_THREADS, sizeof(Smem)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return std::min(per_sm * sms, 8 * OL_P1M / 2);  // the middle system holds 2 rows per CTA in 8 OL_P1M rows
}

inline bool onelaunch_no_placement() {  // DZB_OL_PLACEMENT=0: lay the CTAs out by blockIdx (tuning aid)
  static const bool off = [] {
    const char* e = std::getenv("DZB_OL_PLACEMENT");
    return e && e[0] == '0';
  }();
  return off;
}
// rows of scratch the levels B and C need for an m-row system (per array)
inline size_t onelaunch_rows_b(size_t m) { return ((m + BS_TILE - 1) / BS_TILE) * (size_t)OL_ROWS_B; }
inline size_t onelaunch_rows_c(size_t m) { return ((m + BS_TILE - 1) / BS_TILE) * (size_t)OL_ROWS_C; }

// Where the block scheduler puts the CTAs of a full grid (per device and instantiation, asked once): perm[b] = position along the
// bar of CTA b such that the CTAs sharing an SM are neighbours (2 k, 2 k + 1, ...).  Identity (nullptr) unless every SM holds
// the same number of CTAs; a placement that differs from the probed one costs balance, never correctness — perm is a
// permutation whatever the hardware does.
struct OlPlacement {
  bool tried = false;
  int grid = 0;
  int* d_perm = nullptr;
};
template <int P1C, int MINB>
inline OlPlacement& onelaunch_placement() {
  static OlPlacement p[BS_MAX_DEVICES];
  return p[bigsys_device()];
}
template <int P1C, int MINB>
inline void onelaunch_probe(OlPlacement& pl, OneLaunchArgs& args, int grid, void** params, cudaStream_t st) {
  pl.tried = true;
  int* d_smid = nullptr;
  if (cudaMalloc((void**)&d_smid, grid * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  std::vector<int> smid(grid, -1), perm(grid, -1);
  const OneLaunchArgs saved = args;
  args.probe_smid = d_smid;
  bool ok = cudaLaunchCooperativeKernel((const void*)k_tri_onelaunch<P1C, MINB>, dim3((unsigned)grid), dim3(BS_THREADS), params,
                                        sizeof(OneLaunchSmem<P1C>), st) == cudaSuccess &&
            cudaMemcpyAsync(smid.data(), d_smid, grid * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
            cudaStreamSynchronize(st) == cudaSuccess;
  args = saved;
  cudaFree(d_smid);
  if (!ok) {
    cudaGetLastError();
    return;
  }
  if (!onelaunch_layout(smid.data(), grid, perm.data())) return;  // (ol_layout.h)
  if (cudaMalloc((void**)&pl.d_perm, grid * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(pl.d_perm, perm.data(), grid * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    pl.d_perm = nullptr;
    return;
  }
  pl.grid = grid;
}

template <int P1C, int MINB>
inline int onelaunch_try(OneLaunchArgs& args, size_t tiles, int& cap, void** params, cudaStream_t st, unsigned long long* launches,
                         int* grid_out) {
  if (cap < 0) cap = onelaunch_capacity<P1C, MINB>();
  const size_t grid = std::min<size_t>(tiles, (size_t)std::max(cap, 0));
  *grid_out = (int)grid;
  // level D holds 32 rows per tile (tiles + 1: a ragged last tile counts twice in the kernel's split of the tiles over the CTAs)
  if (grid == 0 || (tiles + 1 + grid - 1) / grid > (size_t)(8 * P1C / OL_ROWS_D)) return -1;
  args.perm = nullptr, args.probe_smid = nullptr;
  if (MINB >= 2 && (int)grid == cap && !onelaunch_no_placement()) {  // a full grid: more than one CTA per SM
    OlPlacement& pl = onelaunch_placement<P1C, MINB>();
    if (!pl.tried) onelaunch_probe<P1C, MINB>(pl, args, (int)grid, params, st);
    if (pl.grid == (int)grid) args.perm = pl.d_perm;
  }
  ++*launches;
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_tri_onelaunch<P1C, MINB>, dim3((unsigned)grid), dim3(BS_THREADS), params,
                                              sizeof(OneLaunchSmem<P1C>), st);
  return e == cudaSuccess ? 0 : bigsys_fail(e, "k_tri_onelaunch");
}

// Launches the kernel on (args.k.p0, args.k.m); returns 0 on success, -1 when the one-launch kernel does not apply
// (no cooperative launch, system too long for the CTA level), 1 on a CUDA error (bigsys_error()).
// caps: per-device capacities of the instantiations, asked once.
inline int onelaunch_run(BigSys* s, OneLaunchArgs args, int (&caps)[2], cudaStream_t st, unsigned long long* launches, int* grid_out) {
  const size_t m = (size_t)args.k.m;
  const size_t tiles = (m + BS_TILE - 1) / BS_TILE;
  // (the rule tables live in shared memory; a last tile of one row would leave its CTA a one-row system)
  if (tiles < 2 || args.rule.m <= 0 || args.rule.m > OL_TAB_CAP || m % BS_TILE == 1) return -1;
  CUtensorMap cx, cs;
  int use_tma = s->tma && !((reinterpret_cast<size_t>(args.k.p0) | reinterpret_cast<size_t>(args.s_ex)) & 15) &&
                bigsys_map(s, args.k.p0, m, &cx) && bigsys_map(s, args.s_ex, m, &cs);
  if (!use_tma) std::memset(&cx, 0, sizeof cx), cs = cx;
  void* params[] = {&cx, &cs, &args, &use_tma};
  int rc = onelaunch_try<128, 2>(args, tiles, caps[0], params, st, launches, grid_out);
  if (rc < 0) rc = onelaunch_try<256, 1>(args, tiles, caps[1], params, st, launches, grid_out);
  return rc;
}

}  // namespace dzb
