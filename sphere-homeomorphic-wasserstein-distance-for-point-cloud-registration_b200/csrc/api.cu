// Library-level entry points of the C ABI (include/shwd.h): version, errors, device query.
#include "common.cuh"

namespace shwd {
static thread_local cudaError_t g_last_cuda = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda = e; }
int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
}  // namespace shwd

extern "C" int shwd_version(void) { return 100; }

extern "C" const char* shwd_error_string(int code) {
  switch (code) {
    case SHWD_OK: return "ok";
    case SHWD_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SHWD_ERR_CUDA: return cudaGetErrorString(shwd::g_last_cuda);
    case SHWD_ERR_WORKSPACE: return "workspace too small or misaligned";
    case SHWD_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" int shwd_device_sm_count(void) { return shwd::sm_count(); }
