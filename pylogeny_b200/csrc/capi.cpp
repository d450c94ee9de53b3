// Library-wide C-ABI pieces: error string, version, device selection, launch counter.
#include <cstring>

#include "common.h"

namespace plg {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace plg

extern "C" const char *plg_last_error(void) { return plg::g_err; }
extern "C" int plg_version(void) { return 100; }
extern "C" int64_t plg_launch_count(void) { return plg::g_launches.load(); }

extern "C" int plg_device_count(int *count) {
  PLG_API_BEGIN
  if (!count) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
  *count = n;
  PLG_API_END
}

extern "C" int plg_set_device(int device) {
  PLG_API_BEGIN
  PLG_CUDA(cudaSetDevice(device));
  PLG_API_END
}
