// mesh2d_kernels.cuh — boundary vertices of a triangle mesh on the device ("next" row f3 of SURVEY §8).
// The reference (mesh/mesh_builder.rs:342-500) counts every undirected edge in a HashMap while it parses, keeps the
// edges seen exactly once, collects their end points in a HashSet and merge-sorts them (:468-480, :623-680).  Here:
//   k_edge_count    one thread per triangle edge: open-addressing hash table keyed by (min, max) of the edge, atomicCAS to
//                   claim a slot, atomicAdd to count
//   k_edge_flag     slots counted once flag their two vertices
//   k_flag_count / k_block_scan / k_flag_scatter   ordered stream compaction of the flags = the ascending, duplicate-free
//                   index list the reference obtains by sorting
// The result is a set, so it does not depend on the table's iteration order, exactly like the reference's.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dzb {

constexpr unsigned long long M2_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long m2_hash(unsigned long long k) {  // splitmix64 finaliser
  k ^= k >> 30, k *= 0xBF58476D1CE4E5B9ull;
  k ^= k >> 27, k *= 0x94D049BB133111EBull;
  return k ^ (k >> 31);
}

// edges of triangle t: (t0,t1), (t0,t2), (t2,t1) as in mesh_builder.rs:417-446
// An index >= n_vert (a caller-supplied triangle list is not trusted) never reaches the table — k_edge_flag would write
// flag[] out of bounds with it — and is reported through *bad instead.
__global__ void __launch_bounds__(256) k_edge_count(const uint32_t* __restrict__ tri, long long n_tri, uint32_t n_vert,
                                                    unsigned long long* __restrict__ keys, unsigned* __restrict__ cnt,
                                                    unsigned long long mask, unsigned long long* __restrict__ bad) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < 3 * n_tri; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / 3;
    const int k = (int)(e - 3 * t);
    const uint32_t p = tri[3 * t + (k == 2 ? 2 : 0)], q = tri[3 * t + (k == 0 ? 1 : (k == 1 ? 2 : 1))];
    if (p >= n_vert || q >= n_vert) {
      atomicAdd(bad, 1ull);
      continue;
    }
    const unsigned long long key = ((unsigned long long)min(p, q) << 32) | (unsigned long long)max(p, q);
    unsigned long long h = m2_hash(key) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&keys[h], M2_EMPTY, key);
      if (prev == M2_EMPTY || prev == key) {
        atomicAdd(&cnt[h], 1u);
        break;
      }
      h = (h + 1) & mask;
    }
  }
}
__global__ void __launch_bounds__(256) k_edge_flag(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt,
                                                   unsigned long long slots, unsigned char* __restrict__ flag) {
  for (unsigned long long h = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; h < slots;
       h += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[h];
    if (key != M2_EMPTY && cnt[h] == 1u) flag[key >> 32] = 1, flag[key & 0xffffffffu] = 1;  // counter != 1 filtered out :470
  }
}
// ordered compaction, 1024 vertices per block
__global__ void __launch_bounds__(256) k_flag_count(const unsigned char* __restrict__ flag, long long n, unsigned* __restrict__ block_cnt) {
  const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (base + j < n && flag[base + j]) ++c;
  __shared__ unsigned s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  if (c) atomicAdd(&s, (unsigned)c);
  __syncthreads();
  if (threadIdx.x == 0) block_cnt[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_block_scan(unsigned* __restrict__ block_cnt, long long n_blocks, unsigned long long* __restrict__ total) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n_blocks; base += 1024) {
    const long long i = base + threadIdx.x;
    const unsigned v = i < n_blocks ? block_cnt[i] : 0u;
    unsigned x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
      if ((threadIdx.x & 31) >= d) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned w = warp_sums[threadIdx.x];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, w, d);
        if (threadIdx.x >= d) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const unsigned before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u) + (x - v);
    if (i < n_blocks) block_cnt[i] = before;  // exclusive prefix
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_flag_scatter(const unsigned char* __restrict__ flag, long long n,
                                                      const unsigned* __restrict__ block_off, uint32_t* __restrict__ out) {
  __shared__ unsigned warp_sums[8];
  const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
  bool f[4];
  unsigned c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) f[j] = base + j < n && flag[base + j], c += f[j];
  unsigned x = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) >= d) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
  __syncthreads();
  unsigned off = block_off[blockIdx.x] + (x - c);
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += warp_sums[w];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (f[j]) out[off++] = (uint32_t)(base + j);
}
// GL vertices of the 2D mesh: [a, b, 0, 0, 0, 1] per vertex (mesh_builder.rs:388-411)
__global__ void __launch_bounds__(256) k_mesh2d_vertices(const double* __restrict__ xy, long long n, double* __restrict__ v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* p = v + 6 * i;
  p[0] = xy[2 * i], p[1] = xy[2 * i + 1], p[2] = 0.0, p[3] = 0.0, p[4] = 0.0, p[5] = 1.0;
}

}  // namespace dzb
