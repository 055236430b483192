// fp64_lat.cu -- B200 micro-measurements used in DESIGN.md: dependent-issue latency of DFMA and of the
// MUFU.RSQ64H / RCP64H seeds, and DFMA throughput per SM sub-partition as a function of the number of
// independent chains (ILP) and resident warps.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void mufu_chain(double* out, long long* cyc, int iters) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = y;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
  long long h[1024];
  const int iters = 2000;
#define RUN(ILP, WARPS)                                                                            \
  {                                                                                                \
    dfma_chain<ILP><<<148, 32 * WARPS>>>(out, cyc, iters, 0.999, 1e-3);                            \
    cudaMemcpy(h, cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);                           \
    double c = (double)h[0] / (iters * 16.0);                                                      \
    printf("DFMA ILP=%d warps/SM=%d: %.2f cycles per round of %d dependent-chain steps -> %.2f cyc/DFMA/warp, " \
           "%.3f DFMA warp-instr/cycle/SMSP\n", ILP, WARPS, c, ILP, c / ILP, (WARPS / 4.0) * ILP / c); \
  }
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1)
  RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(8, 4)
  RUN(1, 8) RUN(2, 8) RUN(4, 8) RUN(8, 8)
  RUN(1, 16) RUN(2, 16) RUN(4, 16)
  mufu_chain<<<148, 32>>>(out, cyc, iters);
  cudaMemcpy(h, cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("MUFU.RSQ64H dependent: %.2f cycles\n", (double)h[0] / (iters * 16.0));
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
