// tridiag_kernels.cuh — tridiagonal solves and the explicit-Euler step of the time-dependent diffusion solver.
//
// FAITHFUL kernels restate the reference operation for operation (one thread per bar, no FMA: the translation unit is
// built with -fmad=false): matrix_solver::solve_by_thomas (matrix_solver/mod.rs:17-48),
// utils::tridiagonal_matrix_vector_multiplication (fem/utils.rs:41-62) and
// DiffussionSolverTimeDependent::solve (diffusion_solver/time_dependent.rs:257-280).
//
// The FAST step (k_td_step_fast) is the B200 design for ensembles: the Thomas factors of the step-invariant mass
// matrix are computed once (k_thomas_factor, faithful recurrences), the per-step operator is pre-scaled,
//     beta_i = A'_lo u_{i-1} + A'_di u_i + A'_up u_{i+1},   A' = (dt S + M) / den_i,
//     d_i    = alpha_i d_{i-1} + beta_i,                    alpha_i = -M_lo,i / den_i,
//     s_i    = d_i - c_i s_{i+1},
// and one warp owns one bar: lane l sweeps rows [l L, (l+1) L) sequentially (Thomas inside the chunk), the 32 chunk
// couplings are resolved by a warp-shuffle scan of the affine maps (the PCR-like step), and a second sweep applies the
// carries.  Operators are stored "chunk-interleaved" (element j of lane l at j*32 + l) so every load is a coalesced
// 256-byte row; u goes through a padded shared-memory transpose.  HBM traffic per DOF-step: u in/out 16 B + A' 24 B +
// alpha 8 B + c 8 B = 56 B (the reference algorithm touches 64 B: u in/out + 3 M + 3 S bands).
#pragma once
#include <cuda_runtime.h>

namespace dzb {

// ---- FAITHFUL: Thomas factorisation of one tridiagonal matrix per thread ------------------------------------------
// c[i] (i < n-1) and den[i] = di_i - lo_i c_{i-1} exactly as solve_by_thomas evaluates them (mod.rs:27-39);
// den[0] = di_0.  c[n-1] := 0 so that back-substitution can treat the last row uniformly.
__global__ void k_thomas_factor(const double* __restrict__ lo, const double* __restrict__ di,
                                const double* __restrict__ up, double* __restrict__ c, double* __restrict__ den,
                                long long n, long long n_bars) {
  const long long bar = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bar >= n_bars) return;
  const long long o = bar * n;
  double cp = up[o] / di[o];
  c[o] = cp;
  den[o] = di[o];
  for (long long i = 1; i < n - 1; ++i) {
    const double dn = di[o + i] - lo[o + i] * cp;
    cp = up[o + i] / dn;
    c[o + i] = cp;
    den[o + i] = dn;
  }
  den[o + n - 1] = di[o + n - 1] - lo[o + n - 1] * cp;
  c[o + n - 1] = 0.0;
}

// ---- FAITHFUL: solve_by_thomas with precomputed (bit-identical) c / den ------------------------------------------
__global__ void k_thomas_solve_faithful(const double* __restrict__ lo, const double* __restrict__ c,
                                        const double* __restrict__ den, const double* __restrict__ rhs,
                                        double* __restrict__ out, long long n, long long n_bars) {
  const long long bar = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bar >= n_bars) return;
  const long long o = bar * n;
  double d = rhs[o] / den[o];
  out[o] = d;
  for (long long i = 1; i < n; ++i) {
    d = (rhs[o + i] - lo[o + i] * d) / den[o + i];
    out[o + i] = d;
  }
  double s = d;
  for (long long i = n - 2; i >= 0; --i) {
    s = out[o + i] - c[o + i] * s;
    out[o + i] = s;
  }
}

// ---- FAITHFUL: one DiffussionSolverTimeDependent::solve(dt) per thread (time_dependent.rs:257-280) ---------------
__global__ void k_td_step_faithful(const double* __restrict__ u, double* __restrict__ u_new,
                                   const double* __restrict__ mlo, const double* __restrict__ mdi,
                                   const double* __restrict__ mup, const double* __restrict__ slo,
                                   const double* __restrict__ sdi, const double* __restrict__ sup,
                                   const double* __restrict__ c, const double* __restrict__ den, long long n,
                                   long long n_bars, double dt, double bc0, double bc1) {
  const long long bar = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bar >= n_bars) return;
  const long long o = bar * n;
  // b = add(c*(S u), 1*(M u))   utils.rs:52-59, :16-30
  double d;
  {
    const double b1 = dt * (sdi[o] * u[o] + sup[o] * u[o + 1]);
    const double b2 = 1.0 * (mdi[o] * u[o] + mup[o] * u[o + 1]);
    d = (b1 + b2) / den[o];
    u_new[o] = d;
  }
  for (long long i = 1; i < n - 1; ++i) {
    const double b1 = dt * (slo[o + i] * u[o + i - 1] + sdi[o + i] * u[o + i] + sup[o + i] * u[o + i + 1]);
    const double b2 = 1.0 * (mlo[o + i] * u[o + i - 1] + mdi[o + i] * u[o + i] + mup[o + i] * u[o + i + 1]);
    d = ((b1 + b2) - mlo[o + i] * d) / den[o + i];
    u_new[o + i] = d;
  }
  {
    const long long i = n - 1;
    const double b1 = dt * (slo[o + i] * u[o + i - 1] + sdi[o + i] * u[o + i]);
    const double b2 = 1.0 * (mlo[o + i] * u[o + i - 1] + mdi[o + i] * u[o + i]);
    d = ((b1 + b2) - mlo[o + i] * d) / den[o + i];
    u_new[o + i] = d;
  }
  double s = d;
  for (long long i = n - 2; i >= 0; --i) {
    s = u_new[o + i] - c[o + i] * s;
    u_new[o + i] = s;
  }
  u_new[o] = bc0;          // :273
  u_new[o + n - 1] = bc1;  // :274
}

// ---- FAST path ----------------------------------------------------------------------------------------------------
// Chunk-interleaved position of row r in a bar padded to 32*L rows.
template <int L>
__device__ __forceinline__ int il_pos(int r) { return (r % L) * 32 + r / L; }

// Builds the per-step operator for time step dt from the natural-layout bands and factors (run when dt changes):
//   A' = (dt*S + M)/den, alpha = -M_lo/den, cc = c      -> chunk-interleaved [n_bars][32 L]; padding rows are zero.
// One CTA per bar; the permutation goes through padded shared memory so that both sides are coalesced.
template <int L>
__global__ void __launch_bounds__(256) k_td_prepare(const double* __restrict__ mlo, const double* __restrict__ mdi,
                                                    const double* __restrict__ mup, const double* __restrict__ slo,
                                                    const double* __restrict__ sdi, const double* __restrict__ sup,
                                                    const double* __restrict__ c, const double* __restrict__ den,
                                                    double* __restrict__ a_lo, double* __restrict__ a_di,
                                                    double* __restrict__ a_up, double* __restrict__ alpha,
                                                    double* __restrict__ cc, int n, long long n_bars, double dt) {
  constexpr int NP = 32 * L;
  constexpr int SP = NP + NP / 32;  // +1 pad per 32 entries of the interleaved order
  __shared__ double sm[5][SP];
  for (long long bar = blockIdx.x; bar < n_bars; bar += gridDim.x) {
    const long long o = bar * (long long)n;
    for (int r = threadIdx.x; r < NP; r += blockDim.x) {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0;
      if (r < n) {
        const double inv = 1.0 / den[o + r];
        v0 = (dt * slo[o + r] + mlo[o + r]) * inv;
        v1 = (dt * sdi[o + r] + mdi[o + r]) * inv;
        v2 = (dt * sup[o + r] + mup[o + r]) * inv;
        v3 = r == 0 ? 0.0 : -mlo[o + r] * inv;
        v4 = c[o + r];
      }
      const int p = il_pos<L>(r);
      const int q = p + p / 32;
      sm[0][q] = v0, sm[1][q] = v1, sm[2][q] = v2, sm[3][q] = v3, sm[4][q] = v4;
    }
    __syncthreads();
    const long long oo = bar * (long long)NP;
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
      const int q = p + p / 32;
      a_lo[oo + p] = sm[0][q], a_di[oo + p] = sm[1][q], a_up[oo + p] = sm[2][q], alpha[oo + p] = sm[3][q],
                cc[oo + p] = sm[4][q];
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// One warp per bar, `steps`==1 step per launch.  WARPS warps per CTA; dynamic shared memory = WARPS * (33 L + 2) doubles.
template <int L, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_td_step_fast(const double* __restrict__ u_old, double* __restrict__ u_new,
                                                             const double* __restrict__ a_lo,
                                                             const double* __restrict__ a_di,
                                                             const double* __restrict__ a_up,
                                                             const double* __restrict__ alpha,
                                                             const double* __restrict__ cc, int n, long long n_bars,
                                                             double bc0, double bc1) {
  constexpr int NP = 32 * L;
  constexpr int SW = NP + 32 + 2;  // one pad per chunk (r + r/L), +1 guard on each side
  extern __shared__ double s_u[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long bar = (long long)blockIdx.x * WARPS + wib;
  if (bar >= n_bars) return;
  double* su = s_u + wib * SW + 1;  // su[-1] and su[pad(NP-1)+1] are the zero guards

  // (1) issue the operator loads first (coalesced 256 B per instruction, streaming)
  const long long ob = bar * (long long)NP + lane;
  double al[L], be[L];
#pragma unroll
  for (int j = 0; j < L; ++j) al[j] = __ldcs(alpha + ob + j * 32);

  // (2) u: natural order -> padded smem (transpose to chunk ownership)
  const double* ub = u_old + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * 32 + lane;
    su[r + r / L] = r < n ? __ldcs(ub + r) : 0.0;
  }
  if (lane == 0) su[-1] = 0.0, su[NP + 32] = 0.0;
  __syncwarp();

  // (3) beta_j = A'_lo u_{r-1} + A'_di u_r + A'_up u_{r+1},  r = lane*L + j
  {
    const int r0 = lane * L;
    double um = lane == 0 ? 0.0 : su[(r0 - 1) + (r0 - 1) / L];
    double u0 = su[r0 + lane];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int rn = r0 + j + 1;
      const double up1 = (j == L - 1) ? (lane == 31 ? 0.0 : su[rn + rn / L]) : su[rn + lane];
      const double lo_ = __ldcs(a_lo + ob + j * 32), di_ = __ldcs(a_di + ob + j * 32), up_ = __ldcs(a_up + ob + j * 32);
      be[j] = fma(up_, up1, fma(di_, u0, lo_ * um));
      um = u0, u0 = up1;
    }
  }

  // (4) forward: chunk-local sweep from a zero carry, warp scan of the affine maps, sweep with the carry
  {
    double e = 0.0, p = 1.0;
#pragma unroll
    for (int j = 0; j < L; ++j) e = fma(al[j], e, be[j]), p *= al[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pe = shfl_up_d(e, d), pp = shfl_up_d(p, d);
      if (lane >= d) e = fma(p, pe, e), p *= pp;
    }
    double carry = shfl_up_d(e, 1);
    if (lane == 0) carry = 0.0;
#pragma unroll
    for (int j = 0; j < L; ++j) carry = fma(al[j], carry, be[j]), be[j] = carry;  // be[] now holds d
  }

  // (5) backward: s_r = d_r - c_r s_{r+1}
#pragma unroll
  for (int j = 0; j < L; ++j) al[j] = -__ldcs(cc + ob + j * 32);
  {
    double e = 0.0, p = 1.0;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) e = fma(al[j], e, be[j]), p *= al[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pe = shfl_down_d(e, d), pp = shfl_down_d(p, d);
      if (lane + d < 32) e = fma(p, pe, e), p *= pp;
    }
    double carry = shfl_down_d(e, 1);
    if (lane == 31) carry = 0.0;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) carry = fma(al[j], carry, be[j]), be[j] = carry;  // be[] now holds s
  }

  // (6) Dirichlet values, transpose back through smem, coalesced streaming store
  __syncwarp();
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const int r = lane * L + j;
    su[r + lane] = be[j];
  }
  __syncwarp();
  double* un = u_new + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * 32 + lane;
    if (r < n) {
      double v = su[r + r / L];
      if (r == 0) v = bc0;
      if (r == n - 1) v = bc1;
      __stcs(un + r, v);
    }
  }
}

// ---- LEAN step: the same algorithm with the operator rebuilt on chip ---------------------------------------------
// For linear hats the off-diagonal entries of M and S depend on ONE number per element, its length h:
//     M_lo,r = M_up,r-1 = h q / 8,        S_lo,r = mu t0 / (2 h) + b t1p / 4,     S_up,r-1 = mu t0 / (2 h) - b t1m / 4
// (q = sum w (1 - x^2), t0 = sum w, t1p/m = sum w (1 +- x) over the rule's used nodes; closed form of
// time_dependent.rs:180-233, see fast_row in assembly_kernels.cuh).  So instead of five pre-scaled operator arrays
// (40 B per row and step) the kernel streams three: h, 1/den and the unscaled diagonal dt S_di + M_di (24 B), and forms
//     A'_lo = (dt S_lo + M_lo)/den,  A'_up = (dt S_up + M_up)/den,  alpha = -M_lo/den,  c = M_up/den
// in registers: one reciprocal and a dozen flops per row on an fp64 pipe that is ~6 % busy in k_td_step_fast.
// HBM traffic 40 B per DOF-step instead of 56.  Rows 0 and n-1 are identity rows (time_dependent.rs:236-239).
// Valid while every interior row passes fast_row's interval checks; k_td_prepare_lean verifies that and raises *bad.
template <int L, int WPB>
__global__ void __launch_bounds__(256) k_td_prepare_lean(const double* __restrict__ mesh, const double* __restrict__ mdi,
                                                         const double* __restrict__ sdi, const double* __restrict__ den,
                                                         double* __restrict__ bdi, double* __restrict__ idn,
                                                         double* __restrict__ hh, int n, long long n_bars, double dt,
                                                         double x_hi, double x_lo, int* __restrict__ bad) {
  constexpr int LANES = 32 * WPB;  // lanes that share one bar; lane g owns rows g L .. g L + L - 1
  constexpr int NP = LANES * L;
  constexpr int SP = NP + NP / 32;
  __shared__ double sm[3][SP];
  for (long long bar = blockIdx.x; bar < n_bars; bar += gridDim.x) {
    const long long o = bar * (long long)n;
    for (int r = threadIdx.x; r < NP; r += blockDim.x) {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0;
      if (r < n) {
        v0 = dt * sdi[o + r] + mdi[o + r];
        v1 = 1.0 / den[o + r];
        if (r + 1 < n) v2 = mesh[o + r + 1] - mesh[o + r];
        if (r >= 1 && r + 1 < n) {  // the interval checks of fast_row (assembly_kernels.cuh)
          const double a = mesh[o + r - 1], c = mesh[o + r], e = mesh[o + r + 1];
          const double cp = (c - a) / 2.0, mp = (c + a) / 2.0, cn = (e - c) / 2.0, mn = (e + c) / 2.0, cs = (e - a) / 2.0,
                       ms = (e + a) / 2.0;
          bool ok = (a < c) && (c < e);
          ok = ok && (cp * x_hi + mp < c) && !(cp * x_lo + mp < a) && (cn * x_hi + mn < e) && !(cn * x_lo + mn < c) &&
               (cs * x_hi + ms < e) && !(cs * x_lo + ms < a);
          if (r >= 2) ok = ok && !(a < mesh[o + r - 2]);
          if (r + 2 < n) ok = ok && !(mesh[o + r + 2] < e);
          if (!ok) *bad = 1;
        }
      }
      const int p = (r % L) * LANES + r / L;
      const int q = p + p / 32;
      sm[0][q] = v0, sm[1][q] = v1, sm[2][q] = v2;
    }
    __syncthreads();
    const long long oo = bar * (long long)NP;
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
      const int q = p + p / 32;
      bdi[oo + p] = sm[0][q], idn[oo + p] = sm[1][q], hh[oo + p] = sm[2][q];
    }
    __syncthreads();
  }
}

struct LeanConsts {
  double dtmu;  // dt * mu * t0 / 2
  double q8;    // q / 8
  double kp;    // dt * b * t1p / 4
  double km;    // dt * b * t1m / 4
};

// One CTA of WPB warps per bar, lane g owns L consecutive rows (n <= 32 WPB L).  All three operator streams are
// loaded into registers before any arithmetic (3 L independent coalesced loads in flight per lane); as row j's
// (1/den, diagonal) die, its (alpha, beta) take their registers, so the recurrence state stays in the register file.
// Two warps per 1024-node bar (L = 16) instead of one (L = 32) halves the registers per thread and doubles the warps
// an SM keeps in flight.  The 32 WPB affine maps are scanned with warp shuffles plus one shared-memory hop per warp.
template <int L, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_td_step_lean(const double* __restrict__ u_old, double* __restrict__ u_new,
                                                           const double* __restrict__ bdi, const double* __restrict__ idn,
                                                           const double* __restrict__ hh, int n, double bc0, double bc1,
                                                           LeanConsts K) {
  constexpr int LANES = 32 * WPB;
  constexpr int NP = LANES * L;
  constexpr int SW = NP + LANES + 2;  // one pad per chunk (r + r/L), +1 guard on each side
  __shared__ double s_u[SW];
  __shared__ double xh[WPB], fe[WPB], fp[WPB], ge[WPB], gp[WPB];  // cross-warp hand-overs
  const int g = threadIdx.x, lane = g & 31, w = g >> 5;
  const long long bar = blockIdx.x;
  double* su = s_u + 1;  // su[-1] and su[pad(NP-1)+1] are the zero guards

  const long long ob = bar * (long long)NP + g;
  double cc[L], al[L], be[L];  // cc: h, then c_r = M_up/den;  al: 1/den, then alpha;  be: diagonal, then beta, d, s
#pragma unroll
  for (int j = 0; j < L; ++j) cc[j] = __ldcs(hh + ob + j * LANES);
#pragma unroll
  for (int j = 0; j < L; ++j) al[j] = __ldcs(idn + ob + j * LANES);
#pragma unroll
  for (int j = 0; j < L; ++j) be[j] = __ldcs(bdi + ob + j * LANES);

  const double* ub = u_old + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * LANES + g;
    su[r + r / L] = r < n ? __ldcs(ub + r) : 0.0;
  }
  if (g == 0) su[-1] = 0.0, su[NP + LANES] = 0.0;
  if (WPB > 1 && lane == 31) xh[w] = cc[L - 1];
  if (WPB > 1) __syncthreads(); else __syncwarp();

  const int r0 = g * L, s0 = r0 + g;  // my first row and its padded position
  {
    double um = g == 0 ? 0.0 : su[(r0 - 1) + (r0 - 1) / L];
    double u0 = su[s0];
    double hl = shfl_up_d(cc[L - 1], 1);  // element to the left of my first row
    if (WPB > 1 && lane == 0 && w > 0) hl = xh[w - 1];
    if (g == 0 || hl == 0.0) hl = 1.0;    // no such element (row 0, padding): its coupling is cut below
    double ml = hl * K.q8;
    double El = fma(K.dtmu, __drcp_rn(hl), ml);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int r = r0 + j, rn = r + 1;
      const double up1 = (j == L - 1) ? (g == LANES - 1 ? 0.0 : su[rn + rn / L]) : su[rn + g];
      const bool interior = r >= 1 && r <= n - 2;
      const double hr = r <= n - 2 ? cc[j] : 1.0;  // element to the right of row r (row 0 hands it on to row 1)
      const double mr = hr * K.q8;
      const double Er = fma(K.dtmu, __drcp_rn(hr), mr);
      const double blo = interior ? El + K.kp : 0.0, bup = interior ? Er - K.km : 0.0;
      const double id = al[j];
      be[j] = fma(bup, up1, fma(be[j], u0, blo * um)) * id;
      al[j] = interior ? -ml * id : 0.0;
      cc[j] = interior ? mr * id : 0.0;
      um = u0, u0 = up1, ml = mr, El = Er;
    }
  }

  // forward: chunk-local sweep from a zero carry, scan of the affine maps over all lanes of the bar, sweep with the carry
  {
    double e = 0.0, p = 1.0;
#pragma unroll
    for (int j = 0; j < L; ++j) e = fma(al[j], e, be[j]), p *= al[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pe = shfl_up_d(e, d), pp = shfl_up_d(p, d);
      if (lane >= d) e = fma(p, pe, e), p *= pp;
    }
    double carry = shfl_up_d(e, 1);
    if (lane == 0) carry = 0.0;
    if (WPB > 1) {
      if (lane == 31) fe[w] = e, fp[w] = p;
      __syncthreads();
      double E = 0.0;  // value entering this warp = composition of the earlier warps applied to 0
      for (int v = 0; v < w; ++v) E = fma(fp[v], E, fe[v]);
      const double pin = shfl_up_d(p, 1);  // product of the maps of the earlier lanes of this warp
      carry = lane == 0 ? E : fma(pin, E, carry);
    }
#pragma unroll
    for (int j = 0; j < L; ++j) carry = fma(al[j], carry, be[j]), be[j] = carry;  // d
  }

  // backward: s_r = d_r - c_r s_{r+1}; the result goes straight into the transpose buffer
  {
    double e = 0.0, p = 1.0;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) e = fma(-cc[j], e, be[j]), p *= -cc[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pe = shfl_down_d(e, d), pp = shfl_down_d(p, d);
      if (lane + d < 32) e = fma(p, pe, e), p *= pp;
    }
    double carry = shfl_down_d(e, 1);
    if (lane == 31) carry = 0.0;
    if (WPB > 1) {
      if (lane == 0) ge[w] = e, gp[w] = p;
      __syncthreads();
      double E = 0.0;  // value entering this warp from the later warps
      for (int v = WPB - 1; v > w; --v) E = fma(gp[v], E, ge[v]);
      const double pin = shfl_down_d(p, 1);
      carry = lane == 31 ? E : fma(pin, E, carry);
    } else {
      __syncwarp();  // every lane is done reading u
    }
#pragma unroll
    for (int j = L - 1; j >= 0; --j) carry = fma(-cc[j], carry, be[j]), su[s0 + j] = carry;
  }
  if (WPB > 1) __syncthreads(); else __syncwarp();
  double* un = u_new + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * LANES + g;
    if (r < n) {
      double v = su[r + r / L];
      if (r == 0) v = bc0;
      if (r == n - 1) v = bc1;
      __stcs(un + r, v);
    }
  }
}

// ---- temporal blocking: `steps` solve(dt) per launch ---------------------------------------------------------------
// dzb_ensemble_step(e, dt, steps) with steps > 1 never shows the intermediate states to anyone, so the bar's operator
// (A' pre-scaled by 1/den, alpha, c: five values per row, 8 rows per lane) and its state stay on chip for the whole
// launch: HBM traffic is 40 B per DOF for ALL the steps together instead of per step.  The per-step work is the same
// arithmetic as k_td_step_lean (beta from the neighbours in shared memory, two scans, Dirichlet ends).  Reported
// separately by bench.py ("multi_step"): the reference-facing solve(dt) contract returns u after every step.
template <int L, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_td_step_lean_multi(const double* __restrict__ u_old, double* __restrict__ u_new,
                                                                 const double* __restrict__ bdi,
                                                                 const double* __restrict__ idn,
                                                                 const double* __restrict__ hh, int n, double bc0,
                                                                 double bc1, LeanConsts K, int steps) {
  constexpr int LANES = 32 * WPB;
  constexpr int NP = LANES * L;
  constexpr int SW = NP + LANES + 2;
  __shared__ double s_u[SW];
  __shared__ double xh[WPB], fe[2][WPB], fp[2][WPB], ge[2][WPB], gp[2][WPB];
  const int g = threadIdx.x, lane = g & 31, w = g >> 5;
  const long long bar = blockIdx.x;
  double* su = s_u + 1;

  const long long ob = bar * (long long)NP + g;
  double cc[L], al[L], blo[L], bdd[L], bup[L], be[L];
#pragma unroll
  for (int j = 0; j < L; ++j) cc[j] = __ldcs(hh + ob + j * LANES);
#pragma unroll
  for (int j = 0; j < L; ++j) al[j] = __ldcs(idn + ob + j * LANES);
#pragma unroll
  for (int j = 0; j < L; ++j) bdd[j] = __ldcs(bdi + ob + j * LANES);
  const double* ub = u_old + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * LANES + g;
    su[r + r / L] = r < n ? __ldcs(ub + r) : 0.0;
  }
  if (g == 0) su[-1] = 0.0, su[NP + LANES] = 0.0;
  if (WPB > 1 && lane == 31) xh[w] = cc[L - 1];
  __syncthreads();

  const int r0 = g * L, s0 = r0 + g;
  {  // the operator, once
    double hl = shfl_up_d(cc[L - 1], 1);
    if (WPB > 1 && lane == 0 && w > 0) hl = xh[w - 1];
    if (g == 0 || hl == 0.0) hl = 1.0;
    double ml = hl * K.q8;
    double El = fma(K.dtmu, __drcp_rn(hl), ml);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int r = r0 + j;
      const bool interior = r >= 1 && r <= n - 2;
      const double hr = r <= n - 2 ? cc[j] : 1.0;
      const double mr = hr * K.q8;
      const double Er = fma(K.dtmu, __drcp_rn(hr), mr);
      const double id = al[j];
      blo[j] = interior ? (El + K.kp) * id : 0.0;
      bup[j] = interior ? (Er - K.km) * id : 0.0;
      bdd[j] = bdd[j] * id;
      al[j] = interior ? -ml * id : 0.0;
      cc[j] = interior ? mr * id : 0.0;
      ml = mr, El = Er;
    }
  }

  for (int it = 0; it < steps; ++it) {
    const int ph = it & 1;  // alternate the hand-over slots so that a fast warp cannot overwrite what a slow one still reads
    {
      double um = g == 0 ? 0.0 : su[(r0 - 1) + (r0 - 1) / L];
      double u0 = su[s0];
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int rn = r0 + j + 1;
        const double up1 = (j == L - 1) ? (g == LANES - 1 ? 0.0 : su[rn + rn / L]) : su[rn + g];
        be[j] = fma(bup[j], up1, fma(bdd[j], u0, blo[j] * um));
        um = u0, u0 = up1;
      }
    }
    {
      double e = 0.0, p = 1.0;
#pragma unroll
      for (int j = 0; j < L; ++j) e = fma(al[j], e, be[j]), p *= al[j];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double pe = shfl_up_d(e, d), pp = shfl_up_d(p, d);
        if (lane >= d) e = fma(p, pe, e), p *= pp;
      }
      double carry = shfl_up_d(e, 1);
      if (lane == 0) carry = 0.0;
      if (WPB > 1) {
        if (lane == 31) fe[ph][w] = e, fp[ph][w] = p;
        __syncthreads();  // also: every lane has read its neighbours' u
        double E = 0.0;
        for (int v = 0; v < w; ++v) E = fma(fp[ph][v], E, fe[ph][v]);
        const double pin = shfl_up_d(p, 1);
        carry = lane == 0 ? E : fma(pin, E, carry);
      }
#pragma unroll
      for (int j = 0; j < L; ++j) carry = fma(al[j], carry, be[j]), be[j] = carry;
    }
    {
      double e = 0.0, p = 1.0;
#pragma unroll
      for (int j = L - 1; j >= 0; --j) e = fma(-cc[j], e, be[j]), p *= -cc[j];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double pe = shfl_down_d(e, d), pp = shfl_down_d(p, d);
        if (lane + d < 32) e = fma(p, pe, e), p *= pp;
      }
      double carry = shfl_down_d(e, 1);
      if (lane == 31) carry = 0.0;
      if (WPB > 1) {
        if (lane == 0) ge[ph][w] = e, gp[ph][w] = p;
        __syncthreads();
        double E = 0.0;
        for (int v = WPB - 1; v > w; --v) E = fma(gp[ph][v], E, ge[ph][v]);
        const double pin = shfl_down_d(p, 1);
        carry = lane == 31 ? E : fma(pin, E, carry);
      } else {
        __syncwarp();
      }
#pragma unroll
      for (int j = L - 1; j >= 0; --j) {
        carry = fma(-cc[j], carry, be[j]);
        const int r = r0 + j;
        su[s0 + j] = r == 0 ? bc0 : (r == n - 1 ? bc1 : carry);  // time_dependent.rs:273-274
      }
    }
    if (WPB > 1) __syncthreads(); else __syncwarp();
  }
  double* un = u_new + bar * (long long)n;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const int r = k * LANES + g;
    if (r < n) __stcs(un + r, su[r + r / L]);
  }
}

}  // namespace dzb
