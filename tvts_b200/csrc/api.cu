// Library-level entry points: version, thread-local error message, launch counter, device info.
#include "common.cuh"
#include "../../include/tvts_b200.h"
#include <atomic>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
}  // namespace

int tvts_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void tvts_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int tvts_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

extern "C" int tvts_version(void) { return TVTS_B200_VERSION; }
extern "C" const char* tvts_last_error(void) { return g_err; }
extern "C" long long tvts_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
