// psqrt_async.cuh -- per-lane prefetch rings through shared memory (cp.async / SASS LDGSTS).
//
// The sweeps are register-bound (255 registers, 8 warps per SM), so a value loaded "one step ahead" into
// registers is sunk by ptxas to just before its use and its DRAM latency lands on the critical path
// (ncu: 15 % of the step loop stalled on that one load).  An asynchronous copy into shared memory needs
// no destination registers: each lane copies the inputs of a later step into its own ring slot,
// commits one group per step and, DEPTH steps later, waits for that group and reads the slot.
//
// Layout [DEPTH][NF][blockDim.x] doubles: field f of slot d of thread t at ((d * NF + f) * blockDim.x + t),
// so a warp's accesses to one field are 32 consecutive doubles (no bank conflicts).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psq {

// Programmatic dependent launch (sm_90+): the kernels of a pass are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs may be scheduled while its predecessor
// in the stream is still draining.  Every kernel calls pdl_entry() before it touches global memory:
// griddepcontrol.wait blocks until the predecessor grid has completed and flushed (a no-op for a plain launch);
// launch_dependents lets the successor's launch overlap this kernel (it waits in its own pdl_entry()).  Because every
// kernel of the chain waits, completion is transitive: kernel i+2 never runs ahead of kernel i.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

template <int NF, int DEPTH>
struct LaneRing {
  static constexpr size_t smem_bytes(int threads) { return (size_t)threads * NF * DEPTH * sizeof(double); }
  double* base;  // this thread's column of the ring
  int stride;    // blockDim.x
  __device__ __forceinline__ LaneRing(unsigned char* smem) : base(reinterpret_cast<double*>(smem) + threadIdx.x), stride(blockDim.x) {}
  // start copying NF doubles src[0], src[fstride], ... into slot d (ok = false: commit an empty group so the
  // group count stays one per step)
  __device__ __forceinline__ void issue(int d, bool ok, const double* src, long long fstride) {
    if (ok) {
#pragma unroll
      for (int f = 0; f < NF; ++f) cp_async8(base + (d * NF + f) * stride, src + f * fstride);
    }
    cp_async_commit();
  }
  // the group committed DEPTH - 1 commits ago (and everything older) has landed
  __device__ __forceinline__ void wait_oldest() { cp_async_wait<DEPTH - 1>(); }
  __device__ __forceinline__ double get(int d, int f) const { return base[(d * NF + f) * stride]; }
};

}  // namespace psq
