// Library-wide state: error text, version, device check, launch counter.
#include "common.cuh"

#include <cstring>

namespace emrt {

std::atomic<int64_t> g_launch_count{0};

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace emrt

extern "C" int emrt_version(void) { return EMRT_ABI_VERSION; }

extern "C" const char* emrt_last_error(void) { return emrt::last_error_buffer(); }

extern "C" int64_t emrt_launch_count(void) { return emrt::g_launch_count.load(); }

extern "C" void emrt_reset_launch_count(void) { emrt::g_launch_count.store(0); }

extern "C" int emrt_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return emrt::set_error(EMRT_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return emrt::set_error(EMRT_ERR_ARCH, "device %d is sm_%d%d; emrt_b200 is built for sm_100a only and has no fallback",
                           dev, major, minor);
  return EMRT_OK;
}
