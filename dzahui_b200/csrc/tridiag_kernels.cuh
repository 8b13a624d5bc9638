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

// ---- FAITHFUL: Thomas factorisation, many bars ------------------------------------------------------------------------
// c[i] (i < n-1) and den[i] = di_i - lo_i c_{i-1} exactly as solve_by_thomas evaluates them (mod.rs:27-39);
// den[0] = di_0.  c[n-1] := 0 so that back-substitution can treat the last row uniformly.
// Coalesced: with a thread per bar on the natural [bar][node] layout the lanes
// of a warp are a whole bar (8 KB at 1024 nodes) apart: every load touches 32 different lines and the kernel ran at ~1.3 TB/s
// (2.0 ms for the 65536 x 1024 ensemble).  Here a warp owns 32 bars and walks them 32 rows at a time: the lanes load each
// bar's 32-row piece of lo / di / up as one 256-byte row into a padded shared-memory tile (cp.async), lane b then runs bar b's
// recurrence over the tile (the same expressions in the same order: bit-identical), leaves c over up and den over di, and
// the tile goes back with coalesced stores.
__device__ __forceinline__ void tf_copy8(double* smem_dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
constexpr int TF_ROWS = 32, TF_WARPS = 4;
constexpr size_t TF_SMEM = (size_t)TF_WARPS * 3 * 32 * (TF_ROWS + 1) * sizeof(double);
__global__ void __launch_bounds__(32 * TF_WARPS) k_thomas_factor_tile(const double* __restrict__ lo, const double* __restrict__ di,
                                                                      const double* __restrict__ up, double* __restrict__ c,
                                                                      double* __restrict__ den, long long n, long long n_bars) {
  extern __shared__ double tf_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double(*L)[TF_ROWS + 1] = reinterpret_cast<double(*)[TF_ROWS + 1]>(tf_smem + (size_t)w * 3 * 32 * (TF_ROWS + 1));
  double(*D)[TF_ROWS + 1] = L + 32;
  double(*U)[TF_ROWS + 1] = D + 32;
  const long long bar0 = ((long long)blockIdx.x * TF_WARPS + w) * 32;
  if (bar0 >= n_bars) return;
  const int nb = (int)min(32LL, n_bars - bar0);
  double cp = 0.0;
  for (long long i0 = 0; i0 < n; i0 += TF_ROWS) {
    const long long row = i0 + lane;
    if (row < n) {  // global -> shared without a register round trip: all 3 nb copies of the lane are in flight at once
#pragma unroll 4
      for (int b = 0; b < nb; ++b) {
        const long long g = (bar0 + b) * n + row;
        tf_copy8(&L[b][lane], lo + g), tf_copy8(&D[b][lane], di + g), tf_copy8(&U[b][lane], up + g);
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (lane < nb) {
      const int rows = (int)min((long long)TF_ROWS, n - i0);
      for (int j = 0; j < rows; ++j) {
        const long long i = i0 + j;
        const double d = D[lane][j], u = U[lane][j];
        if (i == 0) {
          cp = u / d;
          U[lane][j] = cp;
        } else if (i < n - 1) {
          const double dn = d - L[lane][j] * cp;
          cp = u / dn;
          U[lane][j] = cp, D[lane][j] = dn;
        } else {
          D[lane][j] = d - L[lane][j] * cp;
          U[lane][j] = 0.0;
        }
      }
    }
    __syncwarp();
    if (row < n) {
#pragma unroll 8
      for (int b = 0; b < nb; ++b) {
        const long long g = (bar0 + b) * n + row;
        c[g] = U[b][lane], den[g] = D[b][lane];
      }
    }
    __syncwarp();
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

// ---- FAITHFUL, one warp per bar ---------------------------------------------------------------------------------------
// The kernels above run a bar's recurrence in one thread and leave every load on its dependent chain: ~600 cycles
// per row measured on one bar (0.32 ms for the reference's own 1000-node case), ten times what the arithmetic needs.  These
// versions give a bar to a warp: the lanes stream 256-row chunks through shared memory (coalesced; the next chunk is
// already in flight in registers while the current one is swept) and lane 0 runs exactly the same expressions in the same
// order on operands that are one LDS away.  Bit-identical to the kernels above (and to the oracle); used when there are
// few bars (the reference's single-bar solvers), where latency, not throughput, is what a caller sees.
// The forward sweep divides by den_i on its dependent chain (an IEEE division is ~130 cycles: reciprocal seed, Newton
// steps, a slow-path check that is also a scheduling barrier).  den is known from the factorisation, so the factor kernel
// also stores r_i = RN(1 / den_i) (__drcp_rn, off every chain) and the sweep forms the quotient by Markstein's sequence
//     q0 = RN(a r),  e0 = RN(a - den q0),  q1 = RN(q0 + e0 r)      (q1 is a faithful rounding of a / den)
//     e1 = a - den q1  (exact, one fma),   q  = RN(q1 + e1 r)
// and q IS the correctly rounded a / den whenever r is the correctly rounded reciprocal, q1 is faithful and nothing
// under/overflows (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic, section 4.7) — five dependent
// operations, ~40 cycles.  The first correction is there because q0 alone can be 2 ulps off when den's significand is
// close to all ones (tools/check_exact_div.py replays the sequence in exact rational arithmetic on adversarial
// operands: e1 always exact, no mismatch in 2 million cases; tests/test_abi_and_host.py runs a sample).  Outside a generous
// safe range (zero, tiny, huge or non-finite numerator; a chunk whose denominators are not all comfortably normal) the
// sweep falls back to the division itself, so the result has the bits of a / den in every case.
__device__ __noinline__ double exact_div_slow(double a, double den) { return a / den; }  // never speculated: a real call
__device__ __forceinline__ double exact_div(double a, double den, double r, bool den_safe) {
  // the sequence is issued unconditionally and the range test, which needs only `a`, runs beside it instead of in front
  const double q0 = a * r;
  const double e0 = fma(-den, q0, a);
  const double q1 = fma(e0, r, q0);
  const double e1 = fma(-den, q1, a);
  double q = fma(e1, r, q1);
  const double t = fabs(a);
  if (!(den_safe && t > 0x1p-800 && t < 0x1p+800)) {
    // a signed zero over a normal den (homogeneous right-hand sides: the common case of the static problem): RN(a r)
    // already has the right sign, the corrections would turn -0 into +0.  Everything else: the division itself.
    if (den_safe && t == 0.0) q = q0;
    else q = exact_div_slow(a, den);
  }
  return q;
}
__device__ __forceinline__ bool den_in_safe_range(double den) {
  const double t = fabs(den);
  return t > 0x1p-200 && t < 0x1p+200;
}

constexpr int TW_CH = 256;         // rows per staged chunk
constexpr int TW_K = TW_CH / 32;   // rows per lane per chunk
constexpr long long TW_MAX_BARS = 32768;  // measured: 2.6x faster than a thread per bar at 4096 bars x 1024 nodes, 11 % at 32768, equal at 65536

struct WarpStage3 {  // three input streams, chunk by chunk, double-buffered through registers
  double ra[TW_K], rb[TW_K], rc[TW_K];
  __device__ __forceinline__ void fetch(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                        long long base, long long n, int lane) {
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const long long i = base + k * 32 + lane;
      const bool in = i >= 0 && i < n;
      ra[k] = in ? a[i] : 0.0, rb[k] = in ? b[i] : 0.0, rc[k] = in ? c[i] : 0.0;
    }
  }
  __device__ __forceinline__ void commit(double* sa, double* sb, double* sc, int lane) const {
#pragma unroll
    for (int k = 0; k < TW_K; ++k) sa[k * 32 + lane] = ra[k], sb[k * 32 + lane] = rb[k], sc[k * 32 + lane] = rc[k];
  }
};

__global__ void __launch_bounds__(32) k_thomas_factor_warp(const double* __restrict__ lo, const double* __restrict__ di,
                                                           const double* __restrict__ up, double* __restrict__ c,
                                                           double* __restrict__ den, double* __restrict__ rden, long long n) {
  __shared__ double s_lo[TW_CH + 2], s_di[TW_CH + 2], s_up[TW_CH + 2], s_c[TW_CH], s_den[TW_CH], s_r[TW_CH];
  const int lane = threadIdx.x;
  const long long o = (long long)blockIdx.x * n;
  lo += o, di += o, up += o, c += o, den += o, rden += o;
  WarpStage3 st;
  st.fetch(lo, di, up, 0, n, lane);
  double cp = 0.0;
  for (long long base = 0; base < n; base += TW_CH) {
    const int len = (int)min((long long)TW_CH, n - base);
    st.commit(s_lo, s_di, s_up, lane);
    __syncwarp();
    if (base + TW_CH < n) st.fetch(lo, di, up, base + TW_CH, n, lane);
    if (lane == 0) {
      int j = 0;
      if (base == 0) {  // row 0
        cp = s_up[0] / s_di[0];
        s_c[0] = cp, s_den[0] = s_di[0], s_r[0] = __drcp_rn(s_di[0]);
        j = 1;
      }
      const int jend = (base + len == n) ? len - 1 : len;  // the last row of the bar is handled below
      // den_i = di_i - lo_i c'_{i-1}, c'_i = up_i / den_i: the quotient through the correctly rounded reciprocal (which the
      // solve kernels want anyway) and exact_div's correction sequence; operands of row j + 1 are read during row j
      double a = s_lo[j], b = s_di[j], u = s_up[j];
#pragma unroll 4
      for (; j < jend; ++j) {
        const double na = s_lo[j + 1], nb = s_di[j + 1], nu = s_up[j + 1];
        const double dn = b - a * cp;
        const double r = __drcp_rn(dn);
        cp = exact_div(u, dn, r, den_in_safe_range(dn));
        s_c[j] = cp, s_den[j] = dn, s_r[j] = r;
        a = na, b = nb, u = nu;
      }
      if (base + len == n && (n > 1)) {
        const double dn = s_di[len - 1] - s_lo[len - 1] * cp;
        s_den[len - 1] = dn, s_r[len - 1] = __drcp_rn(dn);
        s_c[len - 1] = 0.0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const int j = k * 32 + lane;
      if (j < len) c[base + j] = s_c[j], den[base + j] = s_den[j], rden[base + j] = s_r[j];
    }
    __syncwarp();
  }
}

// forward + backward sweeps of solve_by_thomas (matrix_solver/mod.rs:17-48, with the bit-identical c / den of the factor
// kernel) on staged operands; `chunk_rhs(base, len)` must fill s_rhs[0..len)
// (all lanes) with the bar's right-hand side rows base.. before the chunk is swept
template <class RhsFn>
__device__ __forceinline__ void warp_thomas_sweeps(const double* __restrict__ lo, const double* __restrict__ c,
                                                   const double* __restrict__ den, const double* __restrict__ rden,
                                                   double* __restrict__ out, long long n, int lane, double* s_lo, double* s_den,
                                                   double* s_r, double* s_rhs, double* s_out, RhsFn chunk_rhs) {
  // ---- forward: d_0 = rhs_0 / den_0, d_i = (rhs_i - lo_i d_{i-1}) / den_i
  WarpStage3 st;
  st.fetch(lo, den, rden, 0, n, lane);
  double d = 0.0;
  for (long long base = 0; base < n; base += TW_CH) {
    const int len = (int)min((long long)TW_CH, n - base);
    st.commit(s_lo, s_den, s_r, lane);
    bool safe = true;
#pragma unroll
    for (int k = 0; k < TW_K; ++k) safe = safe && (base + k * 32 + lane >= n || den_in_safe_range(st.rb[k]));
    safe = __all_sync(0xffffffffu, safe);
    chunk_rhs(base, len);
    __syncwarp();
    if (base + TW_CH < n) st.fetch(lo, den, rden, base + TW_CH, n, lane);
    if (lane == 0) {
      int j = 0;
      if (base == 0) {
        d = exact_div(s_rhs[0], s_den[0], s_r[0], safe);
        s_out[0] = d;
        j = 1;
      }
      // the operands of row j + 1 are read while row j is on the chain (the arrays have one spare slot at the end)
      double a = s_lo[j], dn = s_den[j], r = s_r[j], b = s_rhs[j];
#pragma unroll 4
      for (; j < len; ++j) {
        const double na = s_lo[j + 1], ndn = s_den[j + 1], nr = s_r[j + 1], nb = s_rhs[j + 1];
        d = exact_div(b - a * d, dn, r, safe);
        s_out[j] = d;
        a = na, dn = ndn, r = nr, b = nb;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const int j = k * 32 + lane;
      if (j < len) out[base + j] = s_out[j];
    }
    __syncwarp();
  }
  // ---- backward: x_{n-1} = d_{n-1}, x_i = d_i - c_i x_{i+1}; chunks from the top, s_lo / s_den now hold c / d
  double* s_c = s_lo + 1;  // one spare slot in FRONT this time: the pipelined loop reads index -1
  double* s_d = s_den + 1;
  const long long last = ((n - 1) / TW_CH) * TW_CH;
  double r_c[TW_K], r_d[TW_K];
  auto fetch_b = [&](long long base) {
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const long long i = base + k * 32 + lane;
      r_c[k] = i < n ? c[i] : 0.0, r_d[k] = i < n ? out[i] : 0.0;
    }
  };
  fetch_b(last);
  double x = 0.0;
  for (long long base = last; base >= 0; base -= TW_CH) {
    const int len = (int)min((long long)TW_CH, n - base);
#pragma unroll
    for (int k = 0; k < TW_K; ++k) s_c[k * 32 + lane] = r_c[k], s_d[k * 32 + lane] = r_d[k];
    __syncwarp();
    if (base >= TW_CH) fetch_b(base - TW_CH);
    if (lane == 0) {
      int j = len - 1;
      if (base == last) {  // the bar's last row: x = d
        x = s_d[j];
        s_out[j] = x;
        --j;
      }
      double cc = s_c[j], dd = s_d[j];
#pragma unroll 4
      for (; j >= 0; --j) {
        const double nc = s_c[j - 1], nd = s_d[j - 1];
        x = dd - cc * x;
        s_out[j] = x;
        cc = nc, dd = nd;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const int j = k * 32 + lane;
      if (j < len) out[base + j] = s_out[j];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) k_thomas_solve_faithful_warp(const double* __restrict__ lo, const double* __restrict__ c,
                                                                   const double* __restrict__ den, const double* __restrict__ rden,
                                                                   const double* __restrict__ rhs, double* __restrict__ out,
                                                                   long long n) {
  __shared__ double s_lo[TW_CH + 2], s_den[TW_CH + 2], s_r[TW_CH + 2], s_rhs[TW_CH + 2], s_out[TW_CH];
  const int lane = threadIdx.x;
  const long long o = (long long)blockIdx.x * n;
  rhs += o;
  warp_thomas_sweeps(lo + o, c + o, den + o, rden + o, out + o, n, lane, s_lo, s_den, s_r, s_rhs, s_out, [&](long long base, int len) {
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const int j = k * 32 + lane;
      if (j < len) s_rhs[j] = rhs[base + j];
    }
  });
}

// one DiffussionSolverTimeDependent::solve(dt) of one bar per warp: the right-hand side rows are independent of one
// another (same expressions as k_td_step_faithful, evaluated by all lanes), the sweeps are warp_thomas_sweeps
__global__ void __launch_bounds__(32) k_td_step_faithful_warp(const double* __restrict__ u, double* __restrict__ u_new,
                                                              const double* __restrict__ mlo, const double* __restrict__ mdi,
                                                              const double* __restrict__ mup, const double* __restrict__ slo,
                                                              const double* __restrict__ sdi, const double* __restrict__ sup,
                                                              const double* __restrict__ c, const double* __restrict__ den,
                                                              const double* __restrict__ rden, long long n, double dt, double bc0,
                                                              double bc1) {
  __shared__ double s_lo[TW_CH + 2], s_den[TW_CH + 2], s_r[TW_CH + 2], s_rhs[TW_CH + 2], s_out[TW_CH];
  const int lane = threadIdx.x;
  const long long o = (long long)blockIdx.x * n;
  u += o, u_new += o, mlo += o, mdi += o, mup += o, slo += o, sdi += o, sup += o;
  warp_thomas_sweeps(mlo, c + o, den + o, rden + o, u_new, n, lane, s_lo, s_den, s_r, s_rhs, s_out, [&](long long base, int len) {
#pragma unroll
    for (int k = 0; k < TW_K; ++k) {
      const int j = k * 32 + lane;
      if (j < len) {
        const long long i = base + j;
        double b1, b2;  // b = add(c*(S u), 1*(M u))   utils.rs:52-59, :16-30
        if (i == 0) {
          b1 = dt * (sdi[i] * u[i] + sup[i] * u[i + 1]);
          b2 = 1.0 * (mdi[i] * u[i] + mup[i] * u[i + 1]);
        } else if (i == n - 1) {
          b1 = dt * (slo[i] * u[i - 1] + sdi[i] * u[i]);
          b2 = 1.0 * (mlo[i] * u[i - 1] + mdi[i] * u[i]);
        } else {
          b1 = dt * (slo[i] * u[i - 1] + sdi[i] * u[i] + sup[i] * u[i + 1]);
          b2 = 1.0 * (mlo[i] * u[i - 1] + mdi[i] * u[i] + mup[i] * u[i + 1]);
        }
        s_rhs[j] = b1 + b2;
      }
    }
  });
  if (lane == 0) u_new[0] = bc0, u_new[n - 1] = bc1;  // time_dependent.rs:273-274
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
