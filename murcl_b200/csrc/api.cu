// Library-level entry points: version, error reporting, device info.
#include <string>

#include "common.cuh"

namespace murcl {

static thread_local std::string t_last_error;
static thread_local int t_row_order = 0;
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
}

int row_order_descending() { return t_row_order; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace murcl

extern "C" {

int murcl_version(void) { return MURCL_ABI_VERSION; }

const char* murcl_last_error(void) { return murcl::t_last_error.c_str(); }

int64_t murcl_launch_count(void) { return murcl::g_launches.load(); }

int murcl_set_row_order(int descending) {
  const int prev = murcl::t_row_order;
  murcl::t_row_order = descending ? 1 : 0;
  return prev;
}

int murcl_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  MURCL_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  MURCL_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  MURCL_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  MURCL_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (cc_major) *cc_major = b;
  if (cc_minor) *cc_minor = c;
  return MURCL_OK;
}

}  // extern "C"
