// TEST INFRASTRUCTURE ONLY: executes the head-dim-generic attention kernels of tvts_b200/csrc/attention_hd.cu on the CPU SIMT stand-in
// (host_simt.h) over the same grids the real launchers use.
//   g++ -O1 -std=c++20 -pthread -shared -fPIC -DTVTS_HOST_SHIM -I tests/host_kernels harness_attn.cpp
#include "host_simt.h"

namespace hd {
#include "../../tvts_b200/csrc/attention_hd.cu"

template <int HD>
void streamed_fwd(const void* qkv, void* out, float* lse, const int* klen, AttnShape a) {
  simt::launch((unsigned)num_blocks_x(a), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_fwd_kernel<HD>((const bf16*)qkv, (bf16*)out, lse, a, klen); });
}
template <int HD>
void streamed_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, const int* klen, AttnShape a) {
  simt::launch((unsigned)num_blocks_x(a, false), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_bwd_kernel<HD, 0>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a, klen); });
  simt::launch((unsigned)num_blocks_x(a, true), (unsigned)a.H, (unsigned)a.B, kThreads,
               [&] { attn_hd_bwd_kernel<HD, 1>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a, klen); });
}
template <int HD, int NW>
void group_fwd(const void* qkv, void* out, float* lse, AttnShape a) {
  simt::launch((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B, NW * 32,
               [&] { attn_hd_group_fwd_kernel<HD, NW>((const bf16*)qkv, (bf16*)out, lse, a); });
}
template <int HD, int NW>
void group_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, AttnShape a) {
  simt::launch((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B, NW * 32,
               [&] { attn_hd_group_bwd_kernel<HD, NW>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
}

// same kernel selection as hd_launch_fwd / hd_launch_bwd; returns 0 streamed, 1 group-resident (+ CLS launch), 2 / 3 time kernels with 1 / 2 row tiles per slot (+ CLS launch)
template <int HD>
int fwd(const void* qkv, void* out, float* lse, const int* klen, AttnShape a, int group) {
  const int gw = (klen == nullptr && group) ? hd_group_warps(a) : 0;
  const int time_k = (klen == nullptr && group) ? hd_time_tiles(a) : 0;
  if (!gw && !time_k) { streamed_fwd<HD>(qkv, out, lse, klen, a); return 0; }
  if (a.mode != 0) { AttnShape c = a; c.cls_only = 1; streamed_fwd<HD>(qkv, out, lse, nullptr, c); }
  if (time_k) {
    simt::launch((unsigned)((a.n + HD_TW - 1) / HD_TW), (unsigned)a.H, (unsigned)a.B, HD_TW * 32, [&] {
      if (time_k == 1) attn_hd_time_fwd_kernel<HD, 1>((const bf16*)qkv, (bf16*)out, lse, a);
      else attn_hd_time_fwd_kernel<HD, 2>((const bf16*)qkv, (bf16*)out, lse, a);
    });
    return 1 + time_k;
  }
  switch (gw) {
    case 2: group_fwd<HD, 2>(qkv, out, lse, a); break;
    case 3: group_fwd<HD, 3>(qkv, out, lse, a); break;
    case 4: group_fwd<HD, 4>(qkv, out, lse, a); break;
    case 5: group_fwd<HD, 5>(qkv, out, lse, a); break;
    case 6: group_fwd<HD, 6>(qkv, out, lse, a); break;
    default: group_fwd<HD, 7>(qkv, out, lse, a); break;
  }
  return 1;
}
template <int HD>
int bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, const int* klen, AttnShape a, int group) {
  const long long rows = (long long)a.B * a.N * a.H;
  simt::launch((unsigned)((rows + 31) / 32), 1, 1, 256,
               [&] { attn_hd_delta_kernel<HD>((const bf16*)out, (const bf16*)dout, delta, a.B, a.N, a.H); });
  const int gw = (klen == nullptr && group) ? hd_group_warps(a) : 0;
  const int time_k = (klen == nullptr && group) ? hd_time_tiles(a) : 0;
  if (!gw && !time_k) { streamed_bwd<HD>(qkv, dout, lse, delta, dqkv, klen, a); return 0; }
  if (a.mode != 0) { AttnShape c = a; c.cls_only = 1; streamed_bwd<HD>(qkv, dout, lse, delta, dqkv, nullptr, c); }
  if (time_k) {
    simt::launch((unsigned)((a.n + HD_TW - 1) / HD_TW), (unsigned)a.H, (unsigned)a.B, HD_TW * 32, [&] {
      if (time_k == 1) attn_hd_time_bwd_kernel<HD, 1>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a);
      else attn_hd_time_bwd_kernel<HD, 2>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a);
    });
    return 1 + time_k;
  }
  switch (gw) {
    case 2: group_bwd<HD, 2>(qkv, dout, lse, delta, dqkv, a); break;
    case 3: group_bwd<HD, 3>(qkv, dout, lse, delta, dqkv, a); break;
    case 4: group_bwd<HD, 4>(qkv, dout, lse, delta, dqkv, a); break;
    case 5: group_bwd<HD, 5>(qkv, dout, lse, delta, dqkv, a); break;
    case 6: group_bwd<HD, 6>(qkv, dout, lse, delta, dqkv, a); break;
    default: group_bwd<HD, 7>(qkv, dout, lse, delta, dqkv, a); break;
  }
  return 1;
}
}  // namespace hd

extern "C" {

int h_attn_fwd(const void* qkv, void* out, float* lse, const int* klen, long long B, long long N, long long H, long long d, long long mode,
               long long T, long long n, long long causal, float scale, int group) {
  hd::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (d == 64) return hd::fwd<64>(qkv, out, lse, klen, a, group);
  if (d == 80) return hd::fwd<80>(qkv, out, lse, klen, a, group);
  return -1;
}

int h_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, const int* klen, long long B,
               long long N, long long H, long long d, long long mode, long long T, long long n, long long causal, float scale, int group) {
  hd::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (d == 64) return hd::bwd<64>(qkv, out, dout, lse, delta, dqkv, klen, a, group);
  if (d == 80) return hd::bwd<80>(qkv, out, dout, lse, delta, dqkv, klen, a, group);
  return -1;
}

}  // extern "C"
