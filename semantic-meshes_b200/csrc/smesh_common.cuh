// Shared host/device helpers of libsmesh_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/smesh.h"

namespace smesh {

// Per-host-thread error text behind smesh_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int num_sms();
// Opts `fn` in to `smem` bytes of dynamic shared memory on the current device (raise-only, thread-safe) and reports how
// many CTAs of `threads` threads with that much shared memory fit on an SM.
int kernel_blocks_per_sm(const void* fn, int threads, size_t smem, int* blocks_per_sm);

#define SMESH_CUDA_CHECK(expr)                                  \
  do                                                            \
  {                                                             \
    cudaError_t smesh_err_ = (expr);                            \
    if (smesh_err_ != cudaSuccess)                              \
    {                                                           \
      return ::smesh::cuda_fail(smesh_err_, #expr);             \
    }                                                           \
  } while (0)

#define SMESH_LAUNCH_CHECK(name)                                \
  do                                                            \
  {                                                             \
    cudaError_t smesh_err_ = cudaGetLastError();                \
    if (smesh_err_ != cudaSuccess)                              \
    {                                                           \
      return ::smesh::cuda_fail(smesh_err_, "launch of " name); \
    }                                                           \
  } while (0)

static inline size_t align_up(size_t v, size_t a)
{
  return (v + a - 1) / a * a;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers (mbarrier + bulk async copy = the non-tensor TMA path, SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// hint: suspend-time hint of mbarrier.try_wait in ns (the thread may sleep in hardware until the phase completes instead of
// coming back to poll); 0 = plain polling
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t hint = 0x989680u)
{
  if (hint != 0u)
  {
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SMESH_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra SMESH_DONE_%=;\n"
      "bra SMESH_WAIT_%=;\n"
      "SMESH_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(hint)
      : "memory");
  }
  else
  {
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SMESH_WAITP_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SMESH_DONEP_%=;\n"
      "bra SMESH_WAITP_%=;\n"
      "SMESH_DONEP_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
  }
}

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// global -> shared bulk copy, completion signalled on an mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// the same without a cache policy (data that should stay in L2)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global bulk copy in the issuing thread's current bulk group; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_commit_group()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// waits until at most N of the issuing thread's bulk groups have not yet READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ... until at most N have not yet completed entirely
template <int N>
__device__ __forceinline__ void bulk_wait_group()
{
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif // __CUDACC__

} // namespace smesh
