// Error reporting and identification entry points of the C ABI (include/smesh.h).
#include "smesh_common.cuh"

#include <stdarg.h>
#include <string.h>

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

} // namespace smesh

extern "C" const char* smesh_last_error(void)
{
  return smesh::g_error;
}

extern "C" const char* smesh_version(void)
{
  return "smesh_b200 0.1 sm_100a";
}
