// Glue kernels of the TVTS v1 video front end and projection heads (v1/model/video_encoder.py:78-99,178-217,
// v1/model/model_dist_TVTS.py:64-74): tubelet im2col with a PER-TUBE keep mask, token assembly (+ backward), ReLU.
// HBM-bound streaming kernels like elementwise.cu: vectorised coalesced accesses, fp32 arithmetic, exact integer indexing.
#ifdef TVTS_HOST_SHIM          // tests/host_kernels: the kernel bodies below are also compiled for the CPU to execute their index math
#include "host_shim.h"
#else
#include "common.cuh"
#include "../../include/tvts_b200.h"
#endif

namespace {

inline unsigned grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 0x7fffffffLL) g = 0x7fffffffLL;
  return (unsigned)g;
}

// video [B,T,3,R,R] fp32, keep [B, T/2, n] int64 (an independent patch subset per tube, v1/data_loader/YTTemporal_dataset.py:206-215)
// -> cols [(b*nt + tube)*n + j, ((c*2 + dt)*p + u)*p + v] bf16: Conv3d(3, D, k = s = (2,p,p)) on the [B,3,T,H,W] permutation is a
// per-tubelet linear map over (c, dt, u, v), so only the KEPT tubelets are embedded.
__global__ void tubelet_gather_kernel(const float* __restrict__ video, const long long* __restrict__ keep, bf16* __restrict__ cols,
                                      int T, int R, int p, int n, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int pv = p / 4;                 // float4 per patch row
  const int K4 = 3 * 2 * p * pv;        // float4 per output row
  const int nt = T / 2;
  const long long row = i / K4;
  int k4 = (int)(i - row * K4);
  const int j = (int)(row % n);
  const long long bt = row / n;         // b*nt + tube
  const int tube = (int)(bt % nt);
  const long long b = bt / nt;
  const int cd = k4 / (p * pv);         // c*2 + dt
  k4 -= cd * p * pv;
  const int c = cd >> 1, dt = cd & 1;
  const int u = k4 / pv, v4 = k4 - u * pv;
  const int g = R / p;
  const long long pi = keep[bt * n + j];
  const int py = (int)(pi / g), px = (int)(pi % g);
  const long long frame = b * T + 2 * tube + dt;
  const float4 val = *reinterpret_cast<const float4*>(video + ((frame * 3 + c) * R + (py * p + u)) * (long long)R + px * p + v4 * 4);
  reinterpret_cast<uint2*>(cols)[i] = make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
}

// x0[b,0] = cls + pos[0];  x0[b, 1 + t*n + j] = tok[(b*nt+t)*n+j] + pos[1 + keep[b,t,j]] + tem[t]   (video_encoder.py:186-207;
// tok already holds the Conv3d bias)
__global__ void assemble_tube_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                                     const float* __restrict__ tem, const long long* __restrict__ keep, float* __restrict__ x0,
                                     int nt, int n, int D4, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int N = 1 + nt * n;
  const int c = (int)(i % D4);
  const long long rowg = i / D4;
  const int tokpos = (int)(rowg % N);
  const long long b = rowg / N;
  float4 o;
  if (tokpos == 0) {
    const float4 a = reinterpret_cast<const float4*>(cls)[c], q = reinterpret_cast<const float4*>(pos)[c];
    o = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  } else {
    const int t = (tokpos - 1) / n, j = (tokpos - 1) % n;
    const long long src = (b * nt + t) * n + j;
    const long long pi = keep[src];
    const float4 a = reinterpret_cast<const float4*>(tok)[src * D4 + c];
    const float4 q = reinterpret_cast<const float4*>(pos)[(1 + pi) * D4 + c];
    const float4 e = reinterpret_cast<const float4*>(tem)[(long long)t * D4 + c];
    o = make_float4(a.x + q.x + e.x, a.y + q.y + e.y, a.z + q.z + e.z, a.w + q.w + e.w);
  }
  reinterpret_cast<float4*>(x0)[i] = o;
}

// grid (nt + 1, B): block (t, b) handles tube t of sample b (t == nt: the CLS row).  dcls / dpos / dtem are ACCUMULATED.
__global__ void assemble_tube_bwd_kernel(const float* __restrict__ dx0, const long long* __restrict__ keep, float* __restrict__ dcls,
                                         float* __restrict__ dpos, float* __restrict__ dtem, bf16* __restrict__ dtok, int nt, int n,
                                         int D4) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int N = 1 + nt * n;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    if (t == nt) {
      const float4 d = reinterpret_cast<const float4*>(dx0)[((long long)b * N) * D4 + c];
      float* pc = dcls + c * 4;
      float* pp = dpos + c * 4;
      atomicAdd(pc + 0, d.x); atomicAdd(pc + 1, d.y); atomicAdd(pc + 2, d.z); atomicAdd(pc + 3, d.w);
      atomicAdd(pp + 0, d.x); atomicAdd(pp + 1, d.y); atomicAdd(pp + 2, d.z); atomicAdd(pp + 3, d.w);
      continue;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      const long long src = ((long long)b * nt + t) * n + j;
      const float4 d = reinterpret_cast<const float4*>(dx0)[((long long)b * N + 1 + (long long)t * n + j) * D4 + c];
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      const long long pi = keep[src];
      float* pp = dpos + ((1 + pi) * D4 + c) * 4;
      atomicAdd(pp + 0, d.x); atomicAdd(pp + 1, d.y); atomicAdd(pp + 2, d.z); atomicAdd(pp + 3, d.w);
      reinterpret_cast<uint2*>(dtok)[src * D4 + c] = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
    float* pt = dtem + ((long long)t * D4 + c) * 4;
    atomicAdd(pt + 0, acc.x); atomicAdd(pt + 1, acc.y); atomicAdd(pt + 2, acc.z); atomicAdd(pt + 3, acc.w);
  }
}

// y = bf16(max(x, 0))   (txt_proj = Sequential(ReLU, Linear), model_dist_TVTS.py:66-69: the GEMM operand)
__global__ void relu_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)), pack_bf16x2(fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)));
}

// dx = dy * (x > 0)
__global__ void relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const float4 d = reinterpret_cast<const float4*>(dy)[i];
  reinterpret_cast<float4*>(dx)[i] = make_float4(v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f);
}

}  // namespace

#ifndef TVTS_HOST_SHIM
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tvts_tubelet_gather(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p,
                                   int64_t n, void* stream) {
  TVTS_REQUIRE(video && keep_ind && cols, "tubelet_gather: null pointer");
  TVTS_REQUIRE(p > 0 && p % 4 == 0 && R % p == 0, "tubelet_gather: patch=%lld must be a multiple of 4 and divide the resolution", (long long)p);
  TVTS_REQUIRE(T % 2 == 0, "tubelet_gather: T=%lld must be even (tubelets of 2 frames)", (long long)T);
  const long long total = B * (T / 2) * n * 3 * 2 * p * (p / 4);
  if (total == 0) return TVTS_OK;
  tubelet_gather_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(video, (const long long*)keep_ind, (bf16*)cols, (int)T, (int)R, (int)p,
                                                                      (int)n, total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_video_assemble_tube(const float* tok, const float* cls, const float* pos, const float* tem, const int64_t* keep_ind,
                                        float* x0, int64_t B, int64_t nt, int64_t n, int64_t D, void* stream) {
  TVTS_REQUIRE(tok && cls && pos && tem && keep_ind && x0 && D % 4 == 0, "video_assemble_tube: bad arguments");
  const long long total = B * (1 + nt * n) * (D / 4);
  if (total == 0) return TVTS_OK;
  assemble_tube_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(tok, cls, pos, tem, (const long long*)keep_ind, x0, (int)nt, (int)n,
                                                                     (int)(D / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_video_assemble_tube_bwd(const float* dx0, const int64_t* keep_ind, float* dcls, float* dpos, float* dtem,
                                            void* dtok_bf16, int64_t B, int64_t nt, int64_t n, int64_t D, void* stream) {
  TVTS_REQUIRE(dx0 && keep_ind && dcls && dpos && dtem && dtok_bf16 && D % 4 == 0, "video_assemble_tube_bwd: bad arguments");
  if (B == 0) return TVTS_OK;
  TVTS_REQUIRE(B <= 65535, "video_assemble_tube_bwd: grid limits");
  dim3 grid((unsigned)(nt + 1), (unsigned)B);
  assemble_tube_bwd_kernel<<<grid, 192, 0, ST(stream)>>>(dx0, (const long long*)keep_ind, dcls, dpos, dtem, (bf16*)dtok_bf16, (int)nt,
                                                         (int)n, (int)(D / 4));
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_relu_bf16(const float* x, void* y, int64_t n, void* stream) {
  if (n == 0) return TVTS_OK;
  TVTS_REQUIRE(x && y && n > 0 && n % 4 == 0, "relu_bf16: n=%lld must be a positive multiple of 4", (long long)n);
  relu_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(x, (bf16*)y, n / 4);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_relu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
  if (n == 0) return TVTS_OK;
  TVTS_REQUIRE(x && dy && dx && n > 0 && n % 4 == 0, "relu_bwd: n=%lld must be a positive multiple of 4", (long long)n);
  relu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(x, dy, dx, n / 4);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
#endif  // !TVTS_HOST_SHIM
