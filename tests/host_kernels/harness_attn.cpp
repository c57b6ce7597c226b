// TEST INFRASTRUCTURE ONLY: executes the head-dim-generic attention kernels of tvts_b200/csrc/attention_hd.cu on the CPU SIMT stand-in
// (host_simt.h) over the same grids the real launchers use.
//   g++ -O1 -std=c++20 -pthread -shared -fPIC -DTVTS_HOST_SHIM -I tests/host_kernels harness_attn.cpp
#include "host_simt.h"

namespace hd {
#include "../../tvts_b200/csrc/attention_hd.cu"

template <int HD>
void fwd(const void* qkv, void* out, float* lse, const int* klen, AttnShape a) {
  simt::launch((unsigned)num_blocks_x(a), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_fwd_kernel<HD>((const bf16*)qkv, (bf16*)out, lse, a, klen); });
}
template <int HD>
void bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, const int* klen, AttnShape a) {
  const long long rows = (long long)a.B * a.N * a.H;
  simt::launch((unsigned)((rows + 31) / 32), 1, 1, 256,
               [&] { attn_hd_delta_kernel<HD>((const bf16*)out, (const bf16*)dout, delta, a.B, a.N, a.H); });
  simt::launch((unsigned)num_blocks_x(a, false), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_bwd_kernel<HD, 0>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a, klen); });
  simt::launch((unsigned)num_blocks_x(a, true), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_bwd_kernel<HD, 1>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a, klen); });
}
}  // namespace hd

extern "C" {

int h_attn_fwd(const void* qkv, void* out, float* lse, const int* klen, long long B, long long N, long long H, long long d, long long mode,
               long long T, long long n, long long causal, float scale) {
  hd::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (d == 64) hd::fwd<64>(qkv, out, lse, klen, a);
  else if (d == 80) hd::fwd<80>(qkv, out, lse, klen, a);
  else return -1;
  return 0;
}

int h_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, const int* klen, long long B,
               long long N, long long H, long long d, long long mode, long long T, long long n, long long causal, float scale) {
  hd::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (d == 64) hd::bwd<64>(qkv, out, dout, lse, delta, dqkv, klen, a);
  else if (d == 80) hd::bwd<80>(qkv, out, dout, lse, delta, dqkv, klen, a);
  else return -1;
  return 0;
}

}  // extern "C"
