// api.cu — error plumbing and device queries of libmcgm.so.
#include <stdarg.h>
#include "common.cuh"

namespace mcgm {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}

}  // namespace mcgm

extern "C" const char* mcgm_last_error(void) { return mcgm::last_error_ref().c_str(); }
extern "C" int mcgm_version(void) { return 100; }
extern "C" int mcgm_sm_count(void) {
  int n = mcgm::sm_count();
  if (n <= 0) return mcgm::fail(MCGM_E_CUDA, "no CUDA device / cannot query SM count");
  return n;
}
