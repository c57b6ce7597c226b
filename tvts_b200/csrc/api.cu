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
extern "C" int tvts_operand_format(void) { return TVTS_OPERAND_IS_FP16; }
extern "C" const char* tvts_last_error(void) { return g_err; }
extern "C" long long tvts_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// Live per-launch timing of the GEMM kernel (bench.py's roofline line): when enabled, tvts_gemm brackets its launch
// with two CUDA events recorded on the launching stream; tvts_prof_collect sums the elapsed times afterwards.
// ------------------------------------------------------------------------------------------------
namespace {
struct ProfRec { cudaEvent_t e0, e1; double flops; double bytes; long long tag[4]; };
constexpr int kProfMax = 16384;
ProfRec g_prof[kProfMax];
int g_prof_created = 0;
int g_prof_used = 0;
bool g_prof_on = false;
long long g_prof_dropped = 0;
}  // namespace

void tvts_prof_tag(int slot, long long a, long long b, long long c, long long d) {
  if (slot >= 0) { g_prof[slot].tag[0] = a; g_prof[slot].tag[1] = b; g_prof[slot].tag[2] = c; g_prof[slot].tag[3] = d; }
}
bool tvts_prof_begin(cudaStream_t stream, double flops, double bytes, int* slot) {
  *slot = -1;
  if (!g_prof_on) return false;
  if (g_prof_used >= kProfMax) { ++g_prof_dropped; return false; }
  const int i = g_prof_used++;
  if (i >= g_prof_created) {
    cudaEventCreate(&g_prof[i].e0);
    cudaEventCreate(&g_prof[i].e1);
    g_prof_created = i + 1;
  }
  g_prof[i].flops = flops;
  g_prof[i].bytes = bytes;
  cudaEventRecord(g_prof[i].e0, stream);
  *slot = i;
  return true;
}
void tvts_prof_end(cudaStream_t stream, int slot) {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, stream);
}

extern "C" int tvts_prof_enable(int on) {
  g_prof_on = on != 0;
  if (on) { g_prof_used = 0; g_prof_dropped = 0; }
  return TVTS_OK;
}
// Per-record readout (call after synchronising, before tvts_prof_collect): elapsed ms, flops and the 4 tags (M, N, K, flags).
extern "C" int tvts_prof_record(int i, double* ms, double* flops, long long* tags) {
  if (i < 0 || i >= g_prof_used) return TVTS_ERR_INVALID;
  float t = 0.f;
  if (cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1) != cudaSuccess) return TVTS_ERR_CUDA;
  *ms = t; *flops = g_prof[i].flops;
  for (int k = 0; k < 4; ++k) tags[k] = g_prof[i].tag[k];
  return TVTS_OK;
}
extern "C" int tvts_prof_count(void) { return g_prof_used; }
// Call after the stream has been synchronised.  Sums over the recorded launches; returns the number of records.
extern "C" long long tvts_prof_collect(double* total_ms, double* total_flops, double* total_bytes) {
  double ms = 0.0, fl = 0.0, by = 0.0;
  long long n = 0;
  for (int i = 0; i < g_prof_used; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1) != cudaSuccess) continue;
    ms += t; fl += g_prof[i].flops; by += g_prof[i].bytes; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (total_bytes) *total_bytes = by;
  g_prof_used = 0;
  return n;
}
