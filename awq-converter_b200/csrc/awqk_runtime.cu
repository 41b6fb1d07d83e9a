// Host-side runtime glue of the C ABI: error reporting, device guard, version.
#include <cstdio>
#include <cstring>

#include "awqk_common.cuh"

namespace awqk {

static thread_local char g_last_error[512] = "";

void set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  std::snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d: %s", cudaGetErrorName(e),
                cudaGetErrorString(e), file, line, what);
  (void)cudaGetLastError();  // clear the sticky-less error so the next call starts clean
}

DeviceGuard::DeviceGuard(const void* ptr) {
  cudaError_t e = cudaGetDevice(&prev);
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaGetDevice", __FILE__, __LINE__);
    status = AWQK_E_NODEVICE;
    return;
  }
  cur = prev;
  if (ptr != nullptr) {
    cudaPointerAttributes attr;
    e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaPointerGetAttributes", __FILE__, __LINE__);
      status = AWQK_E_CUDA;
      return;
    }
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) {
      std::snprintf(g_last_error, sizeof(g_last_error), "pointer %p is not device memory", ptr);
      status = AWQK_E_BADARG;
      return;
    }
    cur = attr.device;
  }
  if (cur != prev) {
    e = cudaSetDevice(cur);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaSetDevice", __FILE__, __LINE__);
      status = AWQK_E_CUDA;
    }
  }
}

DeviceGuard::~DeviceGuard() {
  if (prev >= 0 && cur != prev) (void)cudaSetDevice(prev);
}

}  // namespace awqk

extern "C" int awqk_version(void) { return AWQK_VERSION; }

extern "C" const char* awqk_error_string(int code) {
  switch (code) {
    case AWQK_OK: return "ok";
    case AWQK_E_BADARG: return "bad argument";
    case AWQK_E_ALIGN: return "misaligned pointer";
    case AWQK_E_CUDA: return "CUDA runtime error";
    case AWQK_E_UNSUPPORTED: return "unsupported combination";
    case AWQK_E_WORKSPACE: return "workspace missing or too small";
    case AWQK_E_NODEVICE: return "no usable CUDA device";
    default: return "unknown error";
  }
}

extern "C" const char* awqk_last_cuda_error(void) { return awqk::g_last_error; }
