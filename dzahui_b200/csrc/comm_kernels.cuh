// comm_kernels.cuh — the SPIKE interface exchange of one bar split across GPUs, device side.
//
// The reference solves the whole system with one serial recurrence (matrix_solver/mod.rs:17-48).  Split into row blocks, one
// per GPU, the only data that has to cross GPUs per solve are the TWO interface rows of every block (8 doubles per rank).
// Two transports carry them:
//   * NCCL (ncclAllGather on the library's stream, dzb_api.cu) — always available;
//   * a peer MAILBOX: every rank owns a small buffer that all other ranks have mapped (cudaIpc*, NVLink peer access).  A
//     rank stores its 8 doubles and then a flag into EVERY rank's mailbox and spins on the flags of its own; nothing is
//     launched for the exchange, so it can sit in the middle of a kernel (k_tri_onelaunch, between its two passes) or be
//     one 32-thread kernel between the reduce and the back-substitution of the materialised path.
// Mailbox layout: [2 slots][MB_MAX ranks][MB_STRIDE doubles]; a record is 8 doubles of payload and a 64-bit flag holding the
// epoch (a per-communicator counter, the same on every rank because every rank makes the same sequence of exchanges).  Slot =
// epoch & 1: a rank can be at most one exchange ahead of the slowest one (it cannot finish exchange e + 1 without that rank's
// record for e + 1, which is written only after the rank has read everything of exchange e), so two slots never collide.
#pragma once
#include <cuda_runtime.h>

namespace dzb {

constexpr int MB_MAX = 8;       // ranks of one box (one process per GPU, 8 x B200 behind NVSwitch)
constexpr int MB_STRIDE = 16;   // doubles per record: 8 payload + flag + padding = one 128-byte line
constexpr size_t MB_DOUBLES = 2 * (size_t)MB_MAX * MB_STRIDE;

struct PeerBox {
  double* box[MB_MAX];          // every rank's mailbox as mapped into this process (box[rank] is the local one)
  int rank, world;              // world <= 1: no exchange
  unsigned long long epoch;     // of this exchange
  int* err;                     // set to 1 when a record did not arrive within timeout_ns (a rank that never launched)
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long mb_now() {
  unsigned long long ns;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
  return ns;
}
__device__ __forceinline__ void mb_store_flag(double* rec, unsigned long long epoch) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(rec + 8)), "l"(epoch) : "memory");
}
__device__ __forceinline__ unsigned long long mb_load_flag(const double* rec) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(reinterpret_cast<const unsigned long long*>(rec + 8)) : "memory");
  return v;
}
__device__ __forceinline__ double mb_load(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// thread `peer` (< world) of one CTA: this rank's 8 doubles -> rank `peer`'s mailbox, then the flag
__device__ __forceinline__ void mb_post(const PeerBox& px, int peer, const double (&v)[8]) {
  double* rec = px.box[peer] + ((size_t)(px.epoch & 1) * MB_MAX + px.rank) * MB_STRIDE;
#pragma unroll
  for (int k = 0; k < 8; ++k) rec[k] = v[k];
  __threadfence_system();
  mb_store_flag(rec, px.epoch);
}
// thread `src` (< world): waits for rank `src`'s record of this epoch in the LOCAL mailbox and copies it to dst[8]
__device__ __forceinline__ void mb_wait(const PeerBox& px, int src, double* dst) {
  const double* rec = px.box[px.rank] + ((size_t)(px.epoch & 1) * MB_MAX + src) * MB_STRIDE;
  const unsigned long long t0 = mb_now();
  unsigned spin = 0;
  while (mb_load_flag(rec) != px.epoch) {
    if ((++spin & 1023u) == 0 && mb_now() - t0 > px.timeout_ns) {  // an error after the time-out, not a hang
      if (px.err) atomicExch(px.err, 1);
      break;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) dst[k] = mb_load(rec + k);
}

// SPIKE interface system, rows[4 i ..] = (a, c, s, f) of unknown i in the order (s_0, e_0, s_1, e_1, ...), m = 2 blocks unknowns.
// Serial Thomas in row-excess form (sigma'_i = s_i + m sigma'_{i-1}, beta_i = sigma'_i - c_i: same-signed sums for an
// M-matrix); every rank runs it redundantly on the gathered rows, so nothing but the 8 doubles per rank crosses NVLink.
// work: 2 m doubles of scratch.
__device__ __forceinline__ void interface_solve_serial(const double* rows, int m, double* ends, double* work) {
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
__global__ void k_interface_solve(const double* __restrict__ rows, int m, double* __restrict__ ends, double* __restrict__ work) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  interface_solve_serial(rows, m, ends, work);
}

// Materialised path over the mailbox, one warp: post this block's 8 doubles (src8), gather everybody's into all[world][8]
// and, with SOLVE, solve the interface system and leave this block's (u_s, u_e) in ends2.  Without SOLVE it is a plain
// all-gather of 8 doubles per rank (the block edge values of a refinement sweep).
template <bool SOLVE>
__global__ void __launch_bounds__(32) k_peer_exchange(const PeerBox px, const double* __restrict__ src8, double* __restrict__ all,
                                                      double* __restrict__ ends2) {
  __shared__ double rows[MB_MAX * 8], ends[2 * MB_MAX], work[4 * MB_MAX];
  const int t = threadIdx.x;
  if (t < px.world) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = src8[k];
    mb_post(px, t, v);
    mb_wait(px, t, rows + 8 * t);
    if (all)
      for (int k = 0; k < 8; ++k) all[8 * t + k] = rows[8 * t + k];
  }
  __syncwarp();
  if (SOLVE && t == 0) {
    interface_solve_serial(rows, 2 * px.world, ends, work);
    ends2[0] = ends[2 * px.rank], ends2[1] = ends[2 * px.rank + 1];
  }
}

}  // namespace dzb
