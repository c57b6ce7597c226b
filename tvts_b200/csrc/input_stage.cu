// GPU-side input stage: the last two steps of the reference's CPU clip transform -- ClipToTensor's x / 255 and
// Normalize's (x - mean[c]) / std[c] (v2/video_transforms/video_transform.py:24-76,627-650, functional.py:81-97; the constants of
// v2/video_transforms/videoaug.py:16) -- fused into the im2col gather of the kept patches.  The data loader then ships uint8 crops
// (4x fewer host->device bytes, no fp32 video tensor in HBM at all).  The three fp32 operations are the reference's, in its order, with
// IEEE round-to-nearest intrinsics (this file must not contract or approximate them), so the bf16 im2col rows are bit-identical to
// patch_gather applied to the normalised fp32 clip.
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

__device__ __forceinline__ float norm_px(unsigned int u, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), stdv);
}

// video [B,T,3,R,R] uint8, keep_ind [B,n] int64 -> cols [(b*T+t)*n + j, c*p*p + u*p + v] bf16   (same mapping as patch_gather_kernel)
__global__ void patch_gather_u8_kernel(const uint8_t* __restrict__ video, const long long* __restrict__ keep, bf16* __restrict__ cols,
                                       int T, int R, int p, int n, float m0, float m1, float m2, float s0, float s1, float s2,
                                       long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int pv = p / 4;                 // 4-pixel groups per patch row
  const int K4 = 3 * p * pv;            // groups per output row
  const long long row = i / K4;
  int k4 = (int)(i - row * K4);
  const int j = (int)(row % n);
  const long long bt = row / n;
  const long long b = bt / T;
  const int c = k4 / (p * pv);
  k4 -= c * p * pv;
  const int u = k4 / pv, v4 = k4 - u * pv;
  const int g = R / p;
  const long long pi = keep[b * n + j];
  const int py = (int)(pi / g), px = (int)(pi % g);
  const uchar4 px4 = *reinterpret_cast<const uchar4*>(video + ((bt * 3 + c) * R + (py * p + u)) * (long long)R + px * p + v4 * 4);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), stdv = c == 0 ? s0 : (c == 1 ? s1 : s2);
  reinterpret_cast<uint2*>(cols)[i] = make_uint2(pack_bf16x2(norm_px(px4.x, mean, stdv), norm_px(px4.y, mean, stdv)),
                                                 pack_bf16x2(norm_px(px4.z, mean, stdv), norm_px(px4.w, mean, stdv)));
}

}  // namespace

extern "C" int tvts_patch_gather_u8(const void* video_u8, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p,
                                    int64_t n, const float* mean3, const float* std3, void* stream) {
  TVTS_REQUIRE(video_u8 && keep_ind && cols && mean3 && std3, "patch_gather_u8: null pointer (mean3 / std3 are HOST arrays of 3 floats)");
  TVTS_REQUIRE(p > 0 && p % 4 == 0 && R % p == 0, "patch_gather_u8: patch=%lld must be a multiple of 4 and divide the resolution", (long long)p);
  TVTS_REQUIRE((uintptr_t)video_u8 % 4 == 0, "patch_gather_u8: video must be 4-byte aligned");
  const long long total = B * T * n * 3 * p * (p / 4);
  if (total == 0) return TVTS_OK;
  const unsigned grid = (unsigned)((total + 255) / 256);
  patch_gather_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint8_t*)video_u8, (const long long*)keep_ind, (bf16*)cols, (int)T, (int)R, (int)p, (int)n, mean3[0], mean3[1], mean3[2], std3[0],
      std3[1], std3[2], total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
