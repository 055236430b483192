// psqrt_tma.cuh -- per-lane record streams staged through shared memory with TMA bulk copies.
//
// Each thread of a sweep walks its own contiguous run of trajectory records (mean [N], factor
// [N][N]) in HBM.  Read or written directly, a warp-wide 8-byte access touches 32 different cache
// lines (32 LSU wavefronts); ncu showed the sweeps stalled on exactly that (lg_throttle).  Here every
// lane instead moves one record at a time between HBM and its private slice of shared memory with
// cp.async.bulk (SASS UBLKCP): one instruction per tile, no registers, no LSU wavefronts, full
// 32-byte sectors on the DRAM side.  The thread then reads / writes its slice with 16-byte shared
// accesses; slices are padded to an odd number of 16-byte words so a quarter-warp hits 32 distinct banks.
//
// Completion: one mbarrier per lane per buffer for loads (expect_tx + try_wait.parity), bulk
// async-groups for stores (commit_group / wait_group.read).  A lane only ever touches its own slice
// and its own barriers, so no cross-lane synchronisation is needed.
//
// Requirements (checked on the host, otherwise the direct-access kernels run): N even (records are
// multiples of 16 bytes) and 16-byte aligned trajectory base pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psq {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, joins the thread's current bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(PENDING) : "memory");
}

// Staging geometry for the trajectory records (mean [N] + factor [N][N] = REC doubles) of state
// dimension N: tiles of S records, NBUF buffers per thread, within ~110 doubles of shared memory
// per thread (227 KB per SM / 256 resident threads).  S = 0 (N odd: records are not multiples of 16
// bytes) selects the direct-access kernels.
template <int N>
struct Cfg {
  static constexpr int REC = N + N * N;
  static constexpr int BUDGET = 110;
  static constexpr int S = (N % 2 != 0) ? 0 : (4 * REC <= BUDGET ? 2 : 1);
  static constexpr int NBUF = (N % 2 != 0) ? 0 : ((S == 2 || 2 * REC <= BUDGET) ? 2 : 1);
  static constexpr int SS = S > 0 ? S : 1, NB = NBUF > 0 ? NBUF : 1;
  static constexpr int BUF_DOUBLES = SS * REC;  // [S][N] means then [S][N][N] factors
  static constexpr int RAW = NB * BUF_DOUBLES;
  static constexpr int LANE = (RAW / 2) % 2 == 1 ? RAW : RAW + 2;  // odd number of 16-byte words
  static constexpr size_t smem_bytes(int threads) { return (size_t)threads * LANE * sizeof(double); }
};

// Odd N: a record (8 N^2 bytes) is not a multiple of 16 bytes, but a PAIR of consecutive records that
// starts at an even trajectory index is (16 N^2 bytes, 16-byte aligned).  The factor stream is staged
// in such pairs; the N-double means and unpaired records at chunk ends are written directly.
template <int N>
struct CfgOdd {
  static constexpr int NBUF = 2;
  static constexpr int PAIR = 2 * N * N;
  static constexpr int RAW = NBUF * PAIR;
  static constexpr int LANE = (RAW / 2) % 2 == 1 ? RAW : RAW + 2;
  static constexpr size_t smem_bytes(int threads) { return (size_t)threads * LANE * sizeof(double); }
};

}  // namespace tma
}  // namespace psq
