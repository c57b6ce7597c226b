// LayerNorm forward / backward.  Reference: fp32 LayerNorm of
// v2/model/video_encoder_ViT_B_16.py:79-85 (eps 1e-5), v2/CLIP/clip/model.py:157-163, and nn.LayerNorm(eps=1e-6)
// of v2/model/sort_transformer.py:73,76,100.
//
// One warp per row, the row held in registers (D/128 float4 per lane, 128-bit coalesced loads), two-pass
// mean / variance like the reference, warp-shuffle reductions.  The forward writes the GEMM A-operand dtype
// (bf16) directly, or fp32 for the residual stream (ln_pre).  The backward fuses: up to two incoming residual
// gradients, the fp32 result, a bf16 copy of the result (operand of the next dgrad/wgrad GEMMs) and the
// dgamma/dbeta column reductions (register partials per warp -> smem -> one atomicAdd per column per CTA).
#ifdef TVTS_HOST_SHIM          // tests/host_kernels: the kernels are also compiled for the CPU SIMT stand-in
#include "host_simt.h"
#else
#include "common.cuh"
#include "../../include/tvts_b200.h"
#endif

namespace {

constexpr int kWarps = 8;

template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(kWarps * 32) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, void* __restrict__ y,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                             long long M, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
    const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
    float4 o;
    o.x = v[i].x * rstd * g.x + b.x; o.y = v[i].y * rstd * g.y + b.y;
    o.z = v[i].z * rstd * g.z + b.z; o.w = v[i].w * rstd * g.w + b.w;
    if (OUT_BF16) {
      reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + row * D)[lane + 32 * i] =
          make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * D)[lane + 32 * i] = o;
    }
  }
}

// Backward: HBM-bound (fp32 x + 16-bit dy + up to two fp32 residual gradients in, fp32 + 16-bit dx out: 11-15 KB per 768-wide row).  A row
// is ONE memory phase: every load of the row (x, dy, res1, res2) is issued before the first reduction, so a warp keeps ~10 KB in flight;
// with the per-column gradient accumulators in registers only one CTA (8 warps) fits an SM, and the round-2 profile of the two-phase
// version (residuals loaded after the warp reductions) showed 37 % of the HBM peak at 12 % occupancy.
template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(kWarps * 32, 1) ln_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                             const float* __restrict__ gamma, const float* __restrict__ res1,
                                                             const float* __restrict__ res2, float* __restrict__ dx,
                                                             bf16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ dxsum, long long M) {
  constexpr int D = NV * 128;
  __shared__ float red[kWarps][128];  // one float4-column group at a time
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 dg[NV], db[NV], dxs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    dxs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < M; row += (long long)gridDim.x * kWarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[NV], gy[NV], rs[NV];
    uint2 dyp[NV];
    // ---- every global load of the row first
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xh[i] = reinterpret_cast<const float4*>(x + row * D)[lane + 32 * i];
      if (DY_BF16) dyp[i] = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + row * D)[lane + 32 * i];
      else gy[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * D)[lane + 32 * i];
    }
    if (res1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) rs[i] = reinterpret_cast<const float4*>(res1 + row * D)[lane + 32 * i];
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (res2) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 r = reinterpret_cast<const float4*>(res2 + row * D)[lane + 32 * i];
        rs[i].x += r.x; rs[i].y += r.y; rs[i].z += r.z; rs[i].w += r.w;
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = xh[i];
      const float4 g4 = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];     // read-only, L1-resident after the first row
      float4 d;
      if (DY_BF16) {
        float2 a = unpack_bf16x2(dyp[i].x), b = unpack_bf16x2(dyp[i].y);
        d = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d = gy[i];
      }
      xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd;
      xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      gy[i].x = d.x * g4.x; gy[i].y = d.y * g4.y; gy[i].z = d.z * g4.z; gy[i].w = d.w * g4.w;
      s1 += gy[i].x + gy[i].y + gy[i].z + gy[i].w;
      s2 += gy[i].x * xh[i].x + gy[i].y * xh[i].y + gy[i].z * xh[i].z + gy[i].w * xh[i].w;
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (gy[i].x - s1 - xh[i].x * s2); o.y = rstd * (gy[i].y - s1 - xh[i].y * s2);
      o.z = rstd * (gy[i].z - s1 - xh[i].z * s2); o.w = rstd * (gy[i].w - s1 - xh[i].w * s2);
      o.x += rs[i].x; o.y += rs[i].y; o.z += rs[i].z; o.w += rs[i].w;
      if (dxsum) { dxs[i].x += o.x; dxs[i].y += o.y; dxs[i].z += o.z; dxs[i].w += o.w; }
      if (dx) reinterpret_cast<float4*>(dx + row * D)[lane + 32 * i] = o;
      if (dx_bf16)
        reinterpret_cast<uint2*>(dx_bf16 + row * D)[lane + 32 * i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
  // column reductions: per float4 group i, columns (lane + 32 i)*4 .. +3
  //   pass 0 dgamma, 1 dbeta, 2 dxsum = column sums of dx (the bias gradient of the Linear whose output gradient dx is)
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    if (dst == nullptr) continue;   // block-uniform
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 v = pass == 0 ? dg[i] : (pass == 1 ? db[i] : dxs[i]);
      __syncthreads();
      *reinterpret_cast<float4*>(&red[warp][lane * 4]) = v;
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
        atomicAdd(dst + i * 128 + threadIdx.x, s);
      }
    }
  }
}

}  // namespace

#ifndef TVTS_HOST_SHIM
#define LN_DISPATCH_NV(D, MACRO)  \
  switch ((D) / 128) {            \
    case 1: MACRO(1); break;      \
    case 2: MACRO(2); break;      \
    case 4: MACRO(4); break;      \
    case 5: MACRO(5); break;      \
    case 6: MACRO(6); break;      \
    case 8: MACRO(8); break;      \
    case 10: MACRO(10); break;    \
    default: return tvts_set_error(TVTS_ERR_UNSUPPORTED, "layernorm: unsupported width D=%lld", (long long)(D)); \
  }

extern "C" int tvts_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int64_t y_is_bf16, float* mean,
                                  float* rstd, int64_t M, int64_t D, float eps, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (M == 0) return TVTS_OK;
  TVTS_REQUIRE(x && gamma && beta && y && M > 0, "layernorm_fwd: bad arguments");
  TVTS_REQUIRE(D % 128 == 0, "layernorm_fwd: D=%lld must be a multiple of 128", (long long)D);
  const unsigned grid = (unsigned)((M + kWarps - 1) / kWarps);
#define LAUNCH(NV_)                                                                                              \
  if (y_is_bf16) ln_fwd_kernel<NV_, true><<<grid, kWarps * 32, 0, st>>>(x, gamma, beta, y, mean, rstd, M, eps); \
  else ln_fwd_kernel<NV_, false><<<grid, kWarps * 32, 0, st>>>(x, gamma, beta, y, mean, rstd, M, eps);
  LN_DISPATCH_NV(D, LAUNCH)
#undef LAUNCH
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_layernorm_bwd_colsum(const void* dy, int64_t dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                         const float* gamma, const float* res1, const float* res2, float* dx, void* dx_bf16,
                                         float* dgamma, float* dbeta, float* dxsum, int64_t M, int64_t D, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (M == 0) return TVTS_OK;
  TVTS_REQUIRE(dy && x && mean && rstd && gamma && M > 0, "layernorm_bwd: bad arguments");
  TVTS_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta must be given together");
  TVTS_REQUIRE(D % 128 == 0, "layernorm_bwd: D=%lld must be a multiple of 128", (long long)D);
  long long want = (M + kWarps - 1) / kWarps;
  const long long cap = 1LL * tvts_num_sms();      // persistent: the register-resident column accumulators allow one CTA per SM
  const unsigned grid = (unsigned)(want < cap ? want : cap);
#define LAUNCH(NV_)                                                                                                          \
  if (dy_is_bf16)                                                                                                            \
    ln_bwd_kernel<NV_, true><<<grid, kWarps * 32, 0, st>>>(dy, x, mean, rstd, gamma, res1, res2, dx, (bf16*)dx_bf16, dgamma, \
                                                           dbeta, dxsum, M);                                                 \
  else                                                                                                                       \
    ln_bwd_kernel<NV_, false><<<grid, kWarps * 32, 0, st>>>(dy, x, mean, rstd, gamma, res1, res2, dx, (bf16*)dx_bf16, dgamma, \
                                                            dbeta, dxsum, M);
  LN_DISPATCH_NV(D, LAUNCH)
#undef LAUNCH
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_layernorm_bwd(const void* dy, int64_t dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* res1, const float* res2, float* dx, void* dx_bf16, float* dgamma,
                                  float* dbeta, int64_t M, int64_t D, void* stream_) {
  return tvts_layernorm_bwd_colsum(dy, dy_is_bf16, x, mean, rstd, gamma, res1, res2, dx, dx_bf16, dgamma, dbeta, nullptr, M, D, stream_);
}
#endif  // !TVTS_HOST_SHIM
