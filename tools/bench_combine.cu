// bench_combine.cu -- latency of ONE level of the sub-warp combines (psqrt_coop2.cuh) as a function of how many
// lane groups share an SM.  Standalone measurement aid (not part of libpsqrt.so):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//        -I sqrt-parallel-smoothers_b200/csrc tools/bench_combine.cu -o tools/bench_combine && tools/bench_combine
// Each CTA keeps IT elements in shared memory and runs `levels` dependent combine levels (item x with item x - 1,
// cyclically) exactly as one Kogge-Stone level of k_mid_scan2 does; time per level = (t(2L) - t(L)) / L.
#include <cstdio>
#include <cuda_runtime.h>

#include "psqrt_coop2.cuh"

using namespace psq;

template <class OP, int IT>
__global__ void __launch_bounds__(IT * OP::G, 1) k_bench(double* out, int levels) {
  constexpr int G = OP::G, NFD = OP::NFD;
  extern __shared__ __align__(16) double sm[];
  double* slots = sm;
  double* wsall = sm + 2 * IT * NFD;
  const int l = threadIdx.x % G, lane = threadIdx.x & 31, gbase = lane - l, x = threadIdx.x / G;
  for (int k = threadIdx.x; k < 2 * IT * NFD; k += blockDim.x) {
    const int off = k % NFD;
    slots[k] = OP::ident(off) * 0.9 + ((off * 7 + k / NFD) % 13) * 1e-3;   // near-identity, finite
  }
  __syncthreads();
  int cur = 0;
  for (int lev = 0; lev < levels; ++lev) {
    OP::combine(slots + (cur * IT + (x + IT - 1) % IT) * NFD, slots + (cur * IT + x) * NFD,
                slots + ((cur ^ 1) * IT + x) * NFD, wsall + x * OP::WS, l, gbase);
    __syncthreads();
    cur ^= 1;
  }
  if (threadIdx.x < NFD) out[blockIdx.x * NFD + threadIdx.x] = slots[cur * IT * NFD + threadIdx.x];
}

template <class OP, int IT>
void run(const char* name, int ctas) {
  const size_t smem = sizeof(double) * (2 * IT * OP::NFD + IT * OP::WS);
  cudaFuncSetAttribute(k_bench<OP, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double* out;
  cudaMalloc(&out, sizeof(double) * ctas * OP::NFD);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float t[2];
  for (int rep = 0; rep < 2; ++rep) {
    const int L = 200 * (rep + 1);
    k_bench<OP, IT><<<ctas, IT * OP::G, smem>>>(out, L);   // warm-up
    cudaEventRecord(e0);
    k_bench<OP, IT><<<ctas, IT * OP::G, smem>>>(out, L);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&t[rep], e0, e1);
  }
  const double us = (t[1] - t[0]) * 1e3 / 200.0;
  printf("%-10s IT=%3d (%4d threads, %2d warps/SM) ctas=%3d : %.3f us / level  (%s)\n", name, IT, IT * OP::G,
         IT * OP::G / 32, ctas, us, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<CoopF2<4>, 4>("filter n=4", 1);
  run<CoopF2<4>, 8>("filter n=4", 1);
  run<CoopF2<4>, 16>("filter n=4", 1);
  run<CoopF2<4>, 32>("filter n=4", 1);
  run<CoopF2<4>, 64>("filter n=4", 1);
  run<CoopF2<4>, 8>("filter n=4", 148);
  run<CoopF2<4>, 64>("filter n=4", 19);
  run<CoopS2<4>, 8>("smooth n=4", 1);
  run<CoopS2<4>, 16>("smooth n=4", 1);
  run<CoopS2<4>, 32>("smooth n=4", 1);
  run<CoopS2<4>, 64>("smooth n=4", 1);
  run<CoopF2<5>, 2>("filter n=5", 1);
  run<CoopF2<5>, 8>("filter n=5", 1);
  run<CoopF2<5>, 32>("filter n=5", 1);
  run<CoopF2<8>, 2>("filter n=8", 1);
  run<CoopF2<8>, 8>("filter n=8", 1);
  run<CoopF2<8>, 32>("filter n=8", 1);
  run<CoopS2<8>, 4>("smooth n=8", 1);
  run<CoopS2<8>, 32>("smooth n=8", 1);
  return 0;
}
