// Losses of the TVTSv2 trainer step, forward and gradient in one call each:
//   InfoNCE on the gathered embeddings = sim_matrix (v2/model/model_dist_TVTSv2_ViT_B_16.py:119-127) followed by
//   NormSoftmaxLoss (v2/model/loss.py:13-25); the gradient is produced for the LOCAL rows only, which is exactly what
//   AllGather_multi.backward keeps (v2/trainer/trainer.py:53-57).
//   sort CE = 2 * nn.CrossEntropyLoss()(pred.reshape(-1, n_trans), labels.reshape(-1))   (v2/trainer/trainer.py:487-492)
// Everything is fp32 on CUDA cores: the problem is tiny (Bg <= a few hundred rows) and precision matters more than
// throughput (the similarity logits are divided by temperature 0.05).
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

// one warp per row: xn = x / max(||x||, eps)
__global__ void normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ norm, int rows, int E,
                                      float eps) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) { const float v = x[(long long)r * E + e]; s += v * v; }
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  const float inv = 1.0f / fmaxf(nrm, eps);
  for (int e = lane; e < E; e += 32) xn[(long long)r * E + e] = x[(long long)r * E + e] * inv;
  if (lane == 0) norm[r] = nrm;
}

// S[i,j] = scale * <an_i, bn_j>.  32x32 output tile per CTA (256 threads, 4 outputs each), fp32 smem tiles.
__global__ void __launch_bounds__(256) sim_kernel(const float* __restrict__ an, const float* __restrict__ bn, float* __restrict__ S,
                                                  int Ra, int Rb, int E, float scale) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int e0 = 0; e0 < E; e0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 8 * k;
      As[r][tx] = (i0 + r < Ra && e0 + tx < E) ? an[(long long)(i0 + r) * E + e0 + tx] : 0.f;
      Bs[r][tx] = (j0 + r < Rb && e0 + tx < E) ? bn[(long long)(j0 + r) * E + e0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float b = Bs[tx][e];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(As[ty + 8 * k][e], b, acc[k]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, j = j0 + tx;
    if (i < Ra && j < Rb) S[(long long)i * Rb + j] = acc[k] * scale;
  }
}

// NormSoftmaxLoss forward on x = S * inv_temp.  warp w < Bg: row w; warp w >= Bg: column w - Bg.
__global__ void nsl_fwd_kernel(const float* __restrict__ S, float* __restrict__ lse_r, float* __restrict__ lse_c,
                               float* __restrict__ loss, int Bg, float inv_temp) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= 2 * Bg) return;
  const bool is_col = w >= Bg;
  const int idx = is_col ? w - Bg : w;
  const long long base = is_col ? idx : (long long)idx * Bg;
  const long long stride = is_col ? Bg : 1;
  float m = -INFINITY;
  for (int k = lane; k < Bg; k += 32) m = fmaxf(m, S[base + k * stride] * inv_temp);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < Bg; k += 32) s += expf(S[base + k * stride] * inv_temp - m);
  s = warp_sum(s);
  const float l = m + logf(s);
  if (lane == 0) {
    (is_col ? lse_c : lse_r)[idx] = l;
    atomicAdd(loss, -(S[(long long)idx * Bg + idx] * inv_temp - l) / Bg);
  }
}

// G[i,j] = dLoss/dS[i,j] = gout * inv_temp * (exp(x_ij - lse_r[i]) + exp(x_ij - lse_c[j]) - 2 delta_ij) / Bg,  x = S * inv_temp
__global__ void nsl_bwd_kernel(const float* __restrict__ S, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                               const float* __restrict__ gout, float* __restrict__ G, int Bg, float inv_temp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Bg * Bg) return;
  const int i = (int)(idx / Bg), j = (int)(idx - (long long)i * Bg);
  const float x = S[idx] * inv_temp;
  const float g = (gout ? *gout : 1.0f) * inv_temp / Bg;
  G[idx] = (expf(x - lse_r[i]) + expf(x - lse_c[j]) - (i == j ? 2.0f : 0.0f)) * g;
}

// Backward of S = scale * normalize(a) normalize(b)^T for one side.  grid (nrows): CTA r handles row (row0 + r) of `self`.
//   transposed = 0: self = a, coefficients G[row, k] (k over rows of other = b)
//   transposed = 1: self = b, coefficients G[k, row] (k over rows of other = a)
__global__ void __launch_bounds__(256) sim_bwd_kernel(const float* __restrict__ G, const float* __restrict__ self_n,
                                                      const float* __restrict__ other_n, const float* __restrict__ self_norm,
                                                      float* __restrict__ dself, int R_self, int R_other, int E, int row0,
                                                      int transposed, float scale, float eps) {
  extern __shared__ float sm[];  // coefficient vector [R_other] + reduction scratch [32]
  float* coef = sm;
  float* red = sm + R_other;
  const int i = row0 + blockIdx.x;
  const int ldg = transposed ? R_self : R_other;  // G is [Ra, Rb] row-major: a-side rows have R_other cols; b-side: G[k, i], ld = Rb = R_self
  for (int k = threadIdx.x; k < R_other; k += blockDim.x)
    coef[k] = (transposed ? G[(long long)k * ldg + i] : G[(long long)i * ldg + k]) * scale;
  __syncthreads();
  const float* sn = self_n + (long long)i * E;
  const float nrm = self_norm[i];
  float dn[4] = {0.f, 0.f, 0.f, 0.f};
  float dotp = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int e = threadIdx.x + c * blockDim.x;
    if (e < E) {
      float a = 0.f;
      for (int k = 0; k < R_other; ++k) a = fmaf(coef[k], other_n[(long long)k * E + e], a);
      dn[c] = a;
      dotp += a * sn[e];
    }
  }
  dotp = warp_sum(dotp);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dotp;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
  // x / max(||x||, eps):  ||x|| > eps -> (dn - xn <xn, dn>) / ||x||;  clamped -> dn / eps
  const bool clamped = nrm <= eps;
  const float inv = 1.0f / fmaxf(nrm, eps);
  float* out = dself + (long long)blockIdx.x * E;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int e = threadIdx.x + c * blockDim.x;
    if (e < E) out[e] = clamped ? dn[c] * inv : (dn[c] - sn[e] * tot) * inv;
  }
}

// one thread per row; C classes (small)
__global__ void sort_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ gout,
                               float* __restrict__ loss, float* __restrict__ dlogits, int R, int C, float weight) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* x = logits + (long long)r * C;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, x[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(x[c] - m);
  const float l = m + logf(s);
  const int y = (int)labels[r];
  if (loss) atomicAdd(loss, weight * (l - x[y]) / R);
  if (dlogits) {
    const float g = (gout ? *gout : 1.0f) * weight / R;
    for (int c = 0; c < C; ++c) dlogits[(long long)r * C + c] = (expf(x[c] - l) - (c == y ? 1.0f : 0.0f)) * g;
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tvts_normalize_rows(const float* x, float* xn, float* norm, int64_t rows, int64_t E, float eps, void* stream) {
  TVTS_REQUIRE(x && xn && norm && E > 0, "normalize_rows: bad arguments");
  if (rows == 0) return TVTS_OK;
  normalize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(x, xn, norm, (int)rows, (int)E, eps);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_sim_matrix(const float* an, const float* bn, float* S, int64_t Ra, int64_t Rb, int64_t E, float scale, void* stream) {
  TVTS_REQUIRE(an && bn && S && E > 0, "sim_matrix: bad arguments");
  if (Ra == 0 || Rb == 0) return TVTS_OK;
  dim3 gs((unsigned)((Rb + 31) / 32), (unsigned)((Ra + 31) / 32));
  sim_kernel<<<gs, 256, 0, ST(stream)>>>(an, bn, S, (int)Ra, (int)Rb, (int)E, scale);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_sim_matrix_bwd(const float* G, const float* self_n, const float* other_n, const float* self_norm, float* dself,
                                   int64_t R_self, int64_t R_other, int64_t E, int64_t row0, int64_t nrows, int64_t transposed,
                                   float scale, float eps, void* stream) {
  TVTS_REQUIRE(G && self_n && other_n && self_norm && dself, "sim_matrix_bwd: null pointer");
  TVTS_REQUIRE(E > 0 && E <= 1024, "sim_matrix_bwd: E=%lld unsupported (<= 1024)", (long long)E);
  TVTS_REQUIRE(row0 >= 0 && row0 + nrows <= R_self, "sim_matrix_bwd: row range out of bounds");
  TVTS_REQUIRE((R_other + 32) * 4 <= 48 * 1024, "sim_matrix_bwd: R_other=%lld too large", (long long)R_other);
  if (nrows == 0) return TVTS_OK;
  sim_bwd_kernel<<<(unsigned)nrows, 256, (R_other + 32) * sizeof(float), ST(stream)>>>(G, self_n, other_n, self_norm, dself, (int)R_self,
                                                                                     (int)R_other, (int)E, (int)row0, (int)transposed,
                                                                                     scale, eps);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_nsl_fwd(const float* S, float* lse_r, float* lse_c, float* loss, int64_t Bg, float temperature, void* stream) {
  TVTS_REQUIRE(S && lse_r && lse_c && loss && Bg > 0 && temperature > 0.f, "nsl_fwd: bad arguments");
  TVTS_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  nsl_fwd_kernel<<<(unsigned)((2 * Bg + 7) / 8), 256, 0, ST(stream)>>>(S, lse_r, lse_c, loss, (int)Bg, 1.0f / temperature);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_nsl_bwd(const float* S, const float* lse_r, const float* lse_c, const float* gout, float* G, int64_t Bg,
                            float temperature, void* stream) {
  TVTS_REQUIRE(S && lse_r && lse_c && G && Bg > 0 && temperature > 0.f, "nsl_bwd: bad arguments");
  nsl_bwd_kernel<<<(unsigned)((Bg * Bg + 255) / 256), 256, 0, ST(stream)>>>(S, lse_r, lse_c, gout, G, (int)Bg, 1.0f / temperature);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_sort_ce(const float* logits, const int64_t* labels, const float* gout, float* loss, float* dlogits, int64_t R,
                            int64_t C, float weight, void* stream) {
  TVTS_REQUIRE(logits && labels && R > 0 && C > 0 && (loss || dlogits), "sort_ce: bad arguments");
  if (loss) TVTS_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  sort_ce_kernel<<<(unsigned)((R + 127) / 128), 128, 0, ST(stream)>>>(logits, (const long long*)labels, gout, loss, dlogits, (int)R,
                                                                      (int)C, weight);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
