// Dependent-issue latency of the fp64 instructions the tridiagonal sweeps chain (one warp, clock64 around an unrolled chain).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 1024;
template <int OP>
__global__ void chain(double x0, double y, double* out, long long* cyc) {
  double x = x0;
  __shared__ double sh[32];
  sh[threadIdx.x] = y;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, y);
    if (OP == 1) x = x + y;
    if (OP == 2) x = x * y;
    if (OP == 3) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
    if (OP == 4) x = __drcp_rn(x);
    if (OP == 5) x = sh[(__double2loint(x) & 31)];  // dependent LDS.64
    if (OP == 6) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = f; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: W warps per SM-quadrant issuing independent DFMAs
__global__ void tput(double y, double* out, long long* cyc) {
  double a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3, e = a + 4, f = a + 5, g = a + 6, h = a + 7;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    a = fma(a, y, y), b = fma(b, y, y), c = fma(c, y, y), d = fma(d, y, y);
    e = fma(e, y, y), f = fma(f, y, y), g = fma(g, y, y), h = fma(h, y, y);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + e + f + g + h;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMallocManaged(&cyc, 8);
  const char* names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H", "__drcp_rn", "LDS.64 dependent", "F2F+FFMA+F2F"};
#define RUN(OP) chain<OP><<<1, 32>>>(1.0000001, 0.9999999, out, cyc); cudaDeviceSynchronize(); chain<OP><<<1, 32>>>(1.0000001, 0.9999999, out, cyc); cudaDeviceSynchronize(); printf("%-18s %.1f cycles/op\n", names[OP], (double)*cyc / N);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
  for (int threads : {32, 128, 256, 512, 1024}) {
    tput<<<1, threads>>>(0.9999999, out, cyc);
    cudaDeviceSynchronize();
    tput<<<1, threads>>>(0.9999999, out, cyc);
    cudaDeviceSynchronize();
    printf("DFMA throughput, %4d threads on one SM: %.2f DFMA lanes/cycle\n", threads, 8.0 * N * threads / (double)*cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
