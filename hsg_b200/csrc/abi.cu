// Error plumbing and device queries behind include/hsg_b200.h.
#include "common.cuh"

#include <mutex>

namespace hsg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cache[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace hsg

extern "C" {

const char* hsg_last_error(void) { return hsg::g_err; }

int hsg_version(void) { return 100; }

int hsg_device_sms(void) {
  int dev = 0, n = 0;
  HSG_CUDA(cudaGetDevice(&dev));
  HSG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

}  // extern "C"
