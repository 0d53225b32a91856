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

} // namespace smesh
