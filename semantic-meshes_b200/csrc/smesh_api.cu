// Error reporting and identification entry points of the C ABI (include/smesh.h).
#include "smesh_common.cuh"

#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

namespace smesh {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what)
{
  set_error("CUDA error %d (%s) in %s", (int) err, cudaGetErrorString(err), what);
  return SMESH_ERR_CUDA;
}

int num_sms()
{
  static thread_local int cached_device = -1;
  static thread_local int cached_sms = 0;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess)
  {
    return 148;
  }
  if (device != cached_device)
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
    {
      sms = 148;
    }
    cached_device = device;
    cached_sms = sms;
  }
  return cached_sms;
}

// Dynamic shared memory opt-in + occupancy of a kernel. cudaFuncAttributeMaxDynamicSharedMemorySize is a property of
// (function, device), shared by every host thread: it is only ever RAISED here, under a lock, so a thread that needs less
// can never pull it below what another thread's next launch asks for.
int kernel_blocks_per_sm(const void* fn, int threads, size_t smem, int* blocks_per_sm)
{
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> raised;                      // (fn, device) -> attribute value
  static std::map<std::tuple<const void*, int, int, size_t>, int> occupancy;        // (fn, device, threads, smem)
  int device = 0;
  SMESH_CUDA_CHECK(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = raised[std::make_pair(fn, device)];
  if (smem > have)
  {
    SMESH_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    have = smem;
  }
  const auto key = std::make_tuple(fn, device, threads, smem);
  auto it = occupancy.find(key);
  if (it == occupancy.end())
  {
    int n = 0;
    SMESH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem));
    it = occupancy.emplace(key, n).first;
  }
  *blocks_per_sm = it->second;
  return SMESH_OK;
}

} // namespace smesh

extern "C" const char* smesh_last_error(void)
{
  return smesh::g_error;
}

extern "C" const char* smesh_version(void)
{
  return "smesh_b200 0.1 sm_100a";
}
