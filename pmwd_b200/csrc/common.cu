#include "common.cuh"

namespace pmwd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace pmwd

extern "C" int pmwd_abi_version(void) { return PMWD_B200_ABI_VERSION; }

extern "C" int pmwd_last_error(char* buf, size_t len) {
  size_t n = strlen(pmwd::g_err);
  if (buf && len) {
    size_t m = n < len - 1 ? n : len - 1;
    memcpy(buf, pmwd::g_err, m);
    buf[m] = 0;
  }
  return (int)n;
}
