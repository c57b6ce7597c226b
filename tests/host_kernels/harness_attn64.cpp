// TEST INFRASTRUCTURE ONLY: executes the SPECIALISED head-dim-64 attention kernels of tvts_b200/csrc/attention.cu (the GPU-verified
// ones: streamed, group-resident, time, CLS) on the CPU SIMT stand-in, with the same kernel selection as tvts_attn_fwd / tvts_attn_bwd
// (without the side stream).  Two purposes: (1) the stand-in itself is checked against kernels whose GPU behaviour is known;
// (2) configurations that were not part of the GPU test list (e.g. causal sequences shorter than 77 tokens) can be checked off-GPU.
//   g++ -O1 -std=c++20 -pthread -shared -fPIC -DTVTS_HOST_SHIM -I tests/host_kernels harness_attn64.cpp
#include "host_simt.h"

namespace a64 {
#include "../../tvts_b200/csrc/attention.cu"

template <int NW>
void group_fwd(const void* qkv, void* out, float* lse, const AttnShape& a) {
  simt::launch((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B, NW * 32,
               [&] { attn_group_fwd_kernel<NW>((const bf16*)qkv, (bf16*)out, lse, a); });
}
template <int NW>
void group_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, const AttnShape& a) {
  simt::launch((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B, NW * 32,
               [&] { attn_group_bwd_kernel<NW>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
}

int fwd(const void* qkv, void* out, float* lse, AttnShape a, int* kinds) {
  const int gw = group_warps(a);
  const bool time_k = use_time_kernels(a);
  const bool split = a.mode != 0 && (time_k || gw);
  if (split) {
    const int smem_bytes = (80 + 8 * 64 + a.N) * 4;
    if (smem_bytes <= 48 * 1024) {
      simt::launch((unsigned)a.H, (unsigned)a.B, 1, CLS_THREADS, [&] { attn_cls_fwd_kernel((const bf16*)qkv, (bf16*)out, lse, a); });
      kinds[0] |= 8;
    } else {
      AttnShape c = a;
      c.cls_only = 1;
      simt::launch((unsigned)num_blocks_x(c), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_fwd_kernel((const bf16*)qkv, (bf16*)out, lse, c); });
    }
  }
  if (time_k) {
    simt::launch((unsigned)((a.n + TW - 1) / TW), (unsigned)a.H, (unsigned)a.B, TW * 32, [&] { attn_time_fwd_kernel((const bf16*)qkv, (bf16*)out, lse, a); });
    kinds[0] |= 4;
  } else if (gw) {
    switch (gw) {
      case 2: group_fwd<2>(qkv, out, lse, a); break;
      case 3: group_fwd<3>(qkv, out, lse, a); break;
      case 4: group_fwd<4>(qkv, out, lse, a); break;
      case 5: group_fwd<5>(qkv, out, lse, a); break;
      case 6: group_fwd<6>(qkv, out, lse, a); break;
      case 7: group_fwd<7>(qkv, out, lse, a); break;
    }
    kinds[0] |= 2;
  } else {
    simt::launch((unsigned)num_blocks_x(a), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_fwd_kernel((const bf16*)qkv, (bf16*)out, lse, a); });
    kinds[0] |= 1;
  }
  return 0;
}

int bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, AttnShape a) {
  const long long rows = (long long)a.B * a.N * a.H;
  simt::launch((unsigned)((rows + 31) / 32), 1, 1, 256, [&] { attn_delta_kernel((const bf16*)out, (const bf16*)dout, delta, a.B, a.N, a.H); });
  const int gw = group_warps(a);
  const bool time_k = use_time_kernels(a);
  const bool split = a.mode != 0 && (time_k || gw);
  if (split) {
    const int smem_bytes = (256 + 8 * 192 + 3 * a.N) * 4;
    if (smem_bytes <= 48 * 1024) {
      simt::launch((unsigned)a.H, (unsigned)a.B, 1, CLS_THREADS, [&] { attn_cls_bwd_kernel((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
    } else {
      AttnShape c = a;
      c.cls_only = 1;
      simt::launch((unsigned)num_blocks_x(c), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_bwd_kernel<0>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, c); });
      simt::launch((unsigned)num_blocks_x(c), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_bwd_kernel<1>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, c); });
    }
  }
  if (time_k) {
    simt::launch((unsigned)((a.n + TW - 1) / TW), (unsigned)a.H, (unsigned)a.B, TW * 32,
                 [&] { attn_time_bwd_kernel((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
  } else if (gw) {
    switch (gw) {
      case 2: group_bwd<2>(qkv, dout, lse, delta, dqkv, a); break;
      case 3: group_bwd<3>(qkv, dout, lse, delta, dqkv, a); break;
      case 4: group_bwd<4>(qkv, dout, lse, delta, dqkv, a); break;
      case 5: group_bwd<5>(qkv, dout, lse, delta, dqkv, a); break;
      case 6: group_bwd<6>(qkv, dout, lse, delta, dqkv, a); break;
      case 7: group_bwd<7>(qkv, dout, lse, delta, dqkv, a); break;
    }
  } else {
    simt::launch((unsigned)num_blocks_x(a), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_bwd_kernel<0>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
    simt::launch((unsigned)num_blocks_x(a), (unsigned)a.H, (unsigned)a.B, kThreads, [&] { attn_bwd_kernel<1>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, a); });
  }
  return 0;
}
}  // namespace a64

extern "C" {
// kinds (out): bit 0 streamed, 1 group-resident, 2 time, 3 CLS kernel -- which forward kernels the dispatch picked
int h64_attn_fwd(const void* qkv, void* out, float* lse, long long B, long long N, long long H, long long mode, long long T, long long n,
                 long long causal, float scale, int* kinds) {
  a64::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  kinds[0] = 0;
  return a64::fwd(qkv, out, lse, a, kinds);
}
int h64_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, long long B, long long N,
                 long long H, long long mode, long long T, long long n, long long causal, float scale) {
  a64::AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  return a64::bwd(qkv, out, dout, lse, delta, dqkv, a);
}
}
