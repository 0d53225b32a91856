// Microbenchmark (GPU box): cost of scattering 80-byte row updates into a [P][20] float accumulator with
//   mode 0: red.global.add.v4.f32, one lane per row (5 instructions, lanes hit different rows)
//   mode 1: red.global.add.v4.f32, 5 adjacent lanes per row (1 instruction covers 6.4 rows)
//   mode 2: cp.reduce.async.bulk (TMA reduce-add) of 80 bytes from shared memory, one op per row
//   mode 3: scalar red.global.add.f32, 20 adjacent lanes per row
// rows are random over P (DRAM-resident accumulator) or confined to a window (L2-resident).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef STRIDE
#define STRIDE 20
#endif

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d)
{
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_f32(float* p, float a)
{
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_mode0(float* acc, const uint32_t* rows, int64_t n)
{
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
  {
    float* dst = acc + (size_t) rows[i] * STRIDE;
#pragma unroll
    for (int j = 0; j < 5; j++) red_v4(dst + 4 * j, 1.f, 2.f, 3.f, 4.f);
  }
}

__global__ void __launch_bounds__(256) k_mode1(float* acc, const uint32_t* rows, int64_t n)
{
  // item = row*5 + chunk
  const int64_t total = n * 5;
  for (int64_t it = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t) gridDim.x * blockDim.x)
  {
    const int64_t r = it / 5;
    const int j = (int) (it - r * 5);
    red_v4(acc + (size_t) rows[r] * STRIDE + 4 * j, 1.f, 2.f, 3.f, 4.f);
  }
}

__global__ void __launch_bounds__(256) k_mode3(float* acc, const uint32_t* rows, int64_t n)
{
  const int64_t total = n * 20;
  for (int64_t it = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t) gridDim.x * blockDim.x)
  {
    const int64_t r = it / 20;
    const int j = (int) (it - r * 20);
    red_f32(acc + (size_t) rows[r] * STRIDE + j, 1.f);
  }
}

__global__ void __launch_bounds__(256) k_mode2(float* acc, const uint32_t* rows, int64_t n)
{
  __shared__ __align__(16) float stage[256 * STRIDE];
  float* mine = stage + threadIdx.x * STRIDE;
  for (int j = 0; j < 20; j++) mine[j] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
  {
    float* dst = acc + (size_t) rows[i] * STRIDE;
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 80;" ::"l"(dst), "r"(smem_u32(mine)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv)
{
  const int64_t P = 2000000, n = 2 * 1024 * 1024;
  float* acc;
  uint32_t* rows;
  cudaMalloc(&acc, P * STRIDE * 4);
  cudaMalloc(&rows, n * 4);
  uint32_t* h = (uint32_t*) malloc(n * 4);
  for (int window = 0; window < 2; window++)
  {
    const int64_t range = window ? 200000 : P; // 16 MB (L2-resident) or 160 MB
    srand(1);
    for (int64_t i = 0; i < n; i++) h[i] = (uint32_t) ((((uint64_t) rand() << 16) ^ rand()) % range);
    cudaMemcpy(rows, h, n * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 4; mode++)
    {
      cudaMemset(acc, 0, P * STRIDE * 4);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float best = 1e9f;
      for (int rep = 0; rep < 5; rep++)
      {
        cudaEventRecord(e0);
        const int grid = 148 * 8;
        if (mode == 0) k_mode0<<<grid, 256>>>(acc, rows, n);
        if (mode == 1) k_mode1<<<grid, 256>>>(acc, rows, n);
        if (mode == 2) k_mode2<<<grid, 256>>>(acc, rows, n);
        if (mode == 3) k_mode3<<<grid, 256>>>(acc, rows, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      cudaError_t err = cudaGetLastError();
      float chk[20];
      cudaMemcpy(chk, acc + (size_t) h[0] * STRIDE, 80, cudaMemcpyDeviceToHost);
      printf("range %8lld rows, mode %d: %8.1f us for %lld row updates (%.2f ns/row, %.1f Mrow/s) err=%s chk=%g %g\n", (long long) range,
             mode, best * 1e3, (long long) n, best * 1e6 / n, n / best / 1e3, cudaGetErrorString(err), chk[0], chk[19]);
    }
  }
  return 0;
}
