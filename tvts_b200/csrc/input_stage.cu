// GPU-side input stage: the last two steps of the reference's CPU clip transform -- ClipToTensor's x / 255 and
// Normalize's (x - mean[c]) / std[c] (v2/video_transforms/video_transform.py:24-76,627-650, functional.py:81-97; the constants of
// v2/video_transforms/videoaug.py:16) -- fused into the im2col gather of the kept patches.  The data loader then ships uint8 crops
// (4x fewer host->device bytes, no fp32 video tensor in HBM at all).  The three fp32 operations are the reference's, in its order, with
// IEEE round-to-nearest intrinsics (this file must not contract or approximate them), so the bf16 im2col rows are bit-identical to
// patch_gather applied to the normalised fp32 clip.
#ifdef TVTS_HOST_SHIM          // tests/host_kernels: the kernel bodies below are also compiled for the CPU to execute their index math
#include "host_shim.h"
#else
#include "common.cuh"
#include "../../include/tvts_b200.h"
#endif

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

#ifndef TVTS_HOST_SHIM
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

#endif  // !TVTS_HOST_SHIM

namespace {
inline unsigned grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 0x7fffffffLL) g = 0x7fffffffLL;
  return (unsigned)g;
}
}  // namespace
#define ST(s) reinterpret_cast<cudaStream_t>(s)

// ================================================================ patch sizes that are not a multiple of 4 / padded GEMM rows (ViT-H/14)
// H/14 has 14x14 patches: K = 3*14*14 = 588 bf16 = 1176 B per im2col row, which is not a multiple of 16 B, so neither TMA nor 16-byte
// vector accesses can walk such rows.  The patch-embed GEMM therefore runs on rows padded to `ld` (a multiple of 8 elements, 592 for
// H/14) whose tail is zero in BOTH operands (exact: the padded products are 0).
namespace {

// same mapping as patch_gather_kernel, two pixels per thread (p even), output row pitch ld, tail [K, ld) zero-filled
__global__ void patch_gather_ld_kernel(const float* __restrict__ video, const long long* __restrict__ keep, bf16* __restrict__ cols,
                                       int B, int T, int R, int p, int n, int ld, long long total_pairs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pairs) return;
  const int L2 = ld / 2;                // bf16 pairs per padded output row
  const int ph = p / 2;                 // pairs per patch row
  const long long row = i / L2;
  int k2 = (int)(i - row * L2);
  uint32_t packed = 0u;
  if (k2 < 3 * p * ph) {
    const int j = (int)(row % n);
    const long long bt = row / n;
    const int b = (int)(bt / T);
    const int c = k2 / (p * ph);
    k2 -= c * p * ph;
    const int u = k2 / ph, v2 = k2 - u * ph;
    const int g = R / p;
    const long long pi = keep[(long long)b * n + j];
    const int py = (int)(pi / g), px = (int)(pi % g);
    const float2 val = *reinterpret_cast<const float2*>(video + ((bt * 3 + c) * R + (py * p + u)) * (long long)R + px * p + v2 * 2);
    packed = pack_bf16x2(val.x, val.y);
  }
  reinterpret_cast<uint32_t*>(cols)[i] = packed;
}

// dst[r, 0:cols] = bf16(src[r, 0:cols]), dst[r, cols:ld] = 0     (src contiguous [rows, cols] fp32; cols, ld even)
__global__ void cast_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cols, int ld, long long total_pairs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pairs) return;
  const int L2 = ld / 2;
  const long long row = i / L2;
  const int k2 = (int)(i - row * L2);
  uint32_t packed = 0u;
  if (2 * k2 < cols) {
    const float2 v = *reinterpret_cast<const float2*>(src + row * cols + 2 * k2);
    packed = pack_bf16x2(v.x, v.y);
  }
  reinterpret_cast<uint32_t*>(dst)[i] = packed;
}

}  // namespace

#ifndef TVTS_HOST_SHIM
extern "C" int tvts_patch_gather_ld(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p,
                                    int64_t n, int64_t ld, void* stream) {
  TVTS_REQUIRE(video && keep_ind && cols, "patch_gather_ld: null pointer");
  TVTS_REQUIRE(p > 0 && p % 2 == 0 && R % p == 0, "patch_gather_ld: patch=%lld must be even and divide the resolution", (long long)p);
  TVTS_REQUIRE(ld % 8 == 0 && ld >= 3 * p * p, "patch_gather_ld: ld=%lld must be a multiple of 8 and >= 3*p*p", (long long)ld);
  const long long total = B * T * n * (ld / 2);
  if (total == 0) return TVTS_OK;
  patch_gather_ld_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(video, (const long long*)keep_ind, (bf16*)cols, (int)B, (int)T,
                                                                       (int)R, (int)p, (int)n, (int)ld, total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_cast_bf16_pad(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld, void* stream) {
  if (rows == 0) return TVTS_OK;
  TVTS_REQUIRE(src && dst && rows > 0, "cast_bf16_pad: bad arguments");
  TVTS_REQUIRE(cols > 0 && cols % 2 == 0 && ld % 2 == 0 && ld >= cols, "cast_bf16_pad: cols=%lld / ld=%lld must be even, ld >= cols",
               (long long)cols, (long long)ld);
  TVTS_REQUIRE((uintptr_t)src % 8 == 0 && (uintptr_t)dst % 4 == 0, "cast_bf16_pad: alignment");
  const long long total = rows * (ld / 2);
  cast_pad_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(src, (bf16*)dst, (int)cols, (int)ld, total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
#endif  // !TVTS_HOST_SHIM
