// ONE launch for everything between the embedding all-gather and the towers' backward passes (SURVEY.md K15 + K16; the north star's
// "fused gathered-logits InfoNCE + sort-CE kernel"):
//   sim_matrix (v2/model/model_dist_TVTSv2_ViT_B_16.py:119-127)  an = a / max(|a|, eps), bn likewise, S = an bn^T  on the GATHERED [Bg, E] embeddings
//   NormSoftmaxLoss (v2/model/loss.py:13-25)                     loss1 = -mean diag log_softmax(S/T, rows) - mean diag log_softmax(S/T, cols)
//   its gradient w.r.t. the LOCAL rows of both embeddings        (AllGather_multi.backward keeps the local slice, v2/trainer/trainer.py:53-57)
//   2 * CrossEntropyLoss()(pred.reshape(-1, C), labels)          (v2/trainer/trainer.py:487-492) and its gradient
// fp32 on CUDA cores: Bg <= 256 rows, E <= 1024 -- the problem is latency-sized (33 MFLOP at Bg = 256), what matters is that it is one
// node of the step graph instead of nine.
//
// One thread-block CLUSTER of up to 8 CTAs.  CTA c owns rows [c * rp, (c+1) * rp) of S (kept in ITS shared memory, later overwritten by
// G = dloss/dS); column statistics and the gradient coefficients G[i, :] / G[:, j] of rows owned by other CTAs are read through
// distributed shared memory; barrier.cluster separates the phases.  Gradients are produced for unit upstream gradients.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kMaxBg = 256;
constexpr int EC = 32;   // embedding columns per staged chunk
constexpr int GT = 8;    // local rows whose gradient is produced per pass of phase 4

struct FusedArgs {
  const float* a;      // video_all [Bg, E]  (rows of S)
  const float* b;      // text_all  [Bg, E]  (columns of S)
  int Bg, E, row0, nloc, rp;
  float inv_temp, eps;
  const float* logits; const long long* labels; int R, C; float ce_weight;
  float* loss1; float* loss2; float* da; float* db; float* dlogits;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < kThreads / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) contrastive_sortce_fused_kernel(FusedArgs p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), cta = (int)cluster.block_rank();
  extern __shared__ float sm[];
  const int Bg = p.Bg, E = p.E, rp = p.rp;
  float* Srow = sm;                         // [rp][Bg]   S, then G
  float* inva = Srow + rp * Bg;             // [Bg] 1 / max(|a_i|, eps)     (negative: clamped row)
  float* invb = inva + Bg;                  // [Bg]
  float* colm = invb + Bg;                  // [Bg] column max over MY rows
  float* cols = colm + Bg;                  // [Bg] column sum of exp(x - colm) over MY rows
  float* lsec = cols + Bg;                  // [Bg]
  float* lser = lsec + Bg;                  // [rp]
  float* coef = lser + rp;                  // [GT][Bg]
  float* red = coef + GT * Bg;              // [32]  (16 warp partials; slot 16 = this CTA's share of loss1)
  float* red2 = red + 32;                   // [16][GT]
  float* As = red2 + (kThreads / 32) * GT;  // [EC][rp + 1]
  float* Bs = As + EC * (rp + 1);           // [EC][Bg + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r_lo = cta * rp, nrows = max(0, min(rp, Bg - r_lo));

  // ---- phase 0: inverse norms of every row of a and b (each CTA for itself: 2 * Bg * E reads out of L2)
  for (int r = warp; r < 2 * Bg; r += kThreads / 32) {
    const float* x = (r < Bg ? p.a : p.b) + (long long)(r < Bg ? r : r - Bg) * E;
    float s = 0.f;
    for (int e = lane; e < E; e += 32) { const float v = x[e]; s += v * v; }
    s = warp_sum(s);
    if (lane == 0) {
      const float nrm = sqrtf(s);
      const float inv = 1.0f / fmaxf(nrm, p.eps);
      (r < Bg ? inva : invb)[r < Bg ? r : r - Bg] = nrm <= p.eps ? -inv : inv;     // sign bit = "clamped" (x / eps branch)
    }
  }
  __syncthreads();

  // ---- phase 1: S[r][j] = <an_r, bn_j> for MY rows, all columns: 4 x 4 outputs per thread, operands staged in chunks of EC columns
  const int tx = tid & 63, ty = tid >> 6;      // 64 column lanes x 8 row lanes
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int e0 = 0; e0 < E; e0 += EC) {
    __syncthreads();
    for (int i = tid; i < rp * EC; i += kThreads) {
      const int r = i / EC, e = i - r * EC;
      As[e * (rp + 1) + r] = (r < nrows && e0 + e < E) ? p.a[(long long)(r_lo + r) * E + e0 + e] * fabsf(inva[r_lo + r]) : 0.f;
    }
    for (int i = tid; i < Bg * EC; i += kThreads) {
      const int j = i / EC, e = i - j * EC;
      Bs[e * (Bg + 1) + j] = (e0 + e < E) ? p.b[(long long)j * E + e0 + e] * fabsf(invb[j]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int e = 0; e < EC; ++e) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = ty + 8 * i; av[i] = r < rp ? As[e * (rp + 1) + r] : 0.f; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = tx + 64 * j; bv[j] = c < Bg ? Bs[e * (Bg + 1) + c] : 0.f; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ty + 8 * i, c = tx + 64 * j;
      if (r < rp && c < Bg) Srow[r * Bg + c] = acc[i][j] * p.inv_temp;       // x = S / T from here on
    }
  __syncthreads();

  // ---- phase 2: row log-sum-exp (my rows), column partials over my rows, then the cluster-wide column log-sum-exp
  for (int r = warp; r < nrows; r += kThreads / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Bg; j += 32) m = fmaxf(m, Srow[r * Bg + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < Bg; j += 32) s += expf(Srow[r * Bg + j] - m);
    s = warp_sum(s);
    if (lane == 0) lser[r] = m + logf(s);
  }
  for (int j = tid; j < Bg; j += kThreads) {
    float m = -INFINITY;
    for (int r = 0; r < nrows; ++r) m = fmaxf(m, Srow[r * Bg + j]);
    float s = 0.f;
    for (int r = 0; r < nrows; ++r) s += expf(Srow[r * Bg + j] - m);
    colm[j] = m; cols[j] = s;
  }
  cluster.sync();
  for (int j = tid; j < Bg; j += kThreads) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, cluster.map_shared_rank(colm, c)[j]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float mc = cluster.map_shared_rank(colm, c)[j];
      if (mc > -INFINITY) s += cluster.map_shared_rank(cols, c)[j] * expf(mc - m);
    }
    lsec[j] = m + logf(s);
  }
  __syncthreads();

  // ---- phase 3: loss1 (my diagonal entries) and G = dloss1/dS in place (unit upstream gradient)
  {
    float part = 0.f;
    for (int r = tid; r < nrows; r += kThreads) {
      const int g = r_lo + r;
      part += -(2.0f * Srow[r * Bg + g] - lser[r] - lsec[g]) / Bg;
    }
    part = block_sum(part, red);
    if (tid == 0) red[kThreads / 32] = part;      // my share of loss1; CTA 0 adds the shares up after the cluster barrier (no atomics, no memset)
    const float gs = p.inv_temp / Bg;
    for (int i = tid; i < nrows * Bg; i += kThreads) {
      const int r = i / Bg, j = i - r * Bg;
      const float x = Srow[i];
      Srow[i] = (expf(x - lser[r]) + expf(x - lsec[j]) - (r_lo + r == j ? 2.0f : 0.0f)) * gs;
    }
  }
  cluster.sync();     // every CTA's G rows (and loss shares) are final
  if (cta == 0 && tid == 0) {
    float l = 0.f;
    for (int c = 0; c < C; ++c) l += cluster.map_shared_rank(red, c)[kThreads / 32];
    *p.loss1 = l;
  }

  // ---- phase 4: gradients of the LOCAL rows.  Local index t -> global index g = row0 + t, handled by CTA t % C, GT of them per pass
  //   side 0 (a / video): coefficients G[g, :] live in the CTA that owns row g;  d an_g = sum_k G[g,k] bn_k
  //   side 1 (b / text):  coefficients G[:, g] are spread over all CTAs;        d bn_g = sum_k G[k,g] an_k
  // (every row of the other side is read ONCE per pass for all GT local rows; one block reduction per pass for the GT dot products)
  for (int side = 0; side < 2; ++side) {
    const float* self = side == 0 ? p.a : p.b;
    const float* other = side == 0 ? p.b : p.a;
    const float* inv_self = side == 0 ? inva : invb;
    const float* inv_other = side == 0 ? invb : inva;
    float* out = side == 0 ? p.da : p.db;
    const int mine = p.nloc > cta ? (p.nloc - cta + C - 1) / C : 0;       // local indices cta, cta + C, ...
    for (int m0 = 0; m0 < mine; m0 += GT) {
      const int ng = min(GT, mine - m0);
      __syncthreads();
      for (int i = tid; i < ng * Bg; i += kThreads) {
        const int q = i / Bg, k = i - q * Bg;
        const int g = p.row0 + cta + (m0 + q) * C;
        float c;
        if (side == 0) c = cluster.map_shared_rank(Srow, g / rp)[(g % rp) * Bg + k];
        else c = cluster.map_shared_rank(Srow, k / rp)[(k % rp) * Bg + g];
        coef[q * Bg + k] = c * fabsf(inv_other[k]);              // folds the other side's normalisation in
      }
      __syncthreads();
      float dn[2][GT];
      float dotp[GT];
#pragma unroll
      for (int q = 0; q < GT; ++q) { dn[0][q] = dn[1][q] = 0.f; dotp[q] = 0.f; }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int e = tid + c * kThreads;
        if (e < E) {
#pragma unroll 8
          for (int k = 0; k < Bg; ++k) {                       // 8 independent row loads in flight (rows come out of L2)
            const float x = other[(long long)k * E + e];
#pragma unroll
            for (int q = 0; q < GT; ++q) dn[c][q] = fmaf(coef[q * Bg + k], x, dn[c][q]);
          }
#pragma unroll
          for (int q = 0; q < GT; ++q)
            if (q < ng) {
              const int g = p.row0 + cta + (m0 + q) * C;
              dotp[q] += dn[c][q] * self[(long long)g * E + e] * fabsf(inv_self[g]);
            }
        }
      }
      // block reduction of the GT dot products: warp sums -> red2[warp][q] -> every thread adds the 16 partials it needs
#pragma unroll
      for (int q = 0; q < GT; ++q) dotp[q] = warp_sum(dotp[q]);
      __syncthreads();
      if (lane == 0)
#pragma unroll
        for (int q = 0; q < GT; ++q) red2[warp * GT + q] = dotp[q];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < GT; ++q) {
        float t2 = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) t2 += red2[w * GT + q];
        dotp[q] = t2;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int e = tid + c * kThreads;
        if (e < E) {
#pragma unroll
          for (int q = 0; q < GT; ++q)
            if (q < ng) {
              const int t = cta + (m0 + q) * C, g = p.row0 + t;
              const float is = inv_self[g];
              const float sn = self[(long long)g * E + e] * fabsf(is);
              out[(long long)t * E + e] = is < 0.f ? dn[c][q] * fabsf(is) : (dn[c][q] - sn * dotp[q]) * fabsf(is);
            }
        }
      }
    }
  }

  // ---- phase 5: sort cross-entropy (CTA 0; one thread per row, C classes)
  if (cta == 0 && p.logits != nullptr) {
    float part = 0.f;
    for (int r = tid; r < p.R; r += kThreads) {
      const float* x = p.logits + (long long)r * p.C;
      float m = -INFINITY;
      for (int c = 0; c < p.C; ++c) m = fmaxf(m, x[c]);
      float s = 0.f;
      for (int c = 0; c < p.C; ++c) s += expf(x[c] - m);
      const float l = m + logf(s);
      const int y = (int)p.labels[r];
      part += p.ce_weight * (l - x[y]) / p.R;
      if (p.dlogits) {
        const float gsc = p.ce_weight / p.R;
        for (int c = 0; c < p.C; ++c) p.dlogits[(long long)r * p.C + c] = (expf(x[c] - l) - (c == y ? 1.0f : 0.0f)) * gsc;
      }
    }
    part = block_sum(part, red);
    if (tid == 0) *p.loss2 = part;
  }
  cluster.sync();     // nobody leaves while its shared memory may still be read remotely
}

size_t fused_smem_bytes(int Bg, int rp) {
  return sizeof(float) * ((size_t)rp * Bg + (5 + GT) * (size_t)Bg + rp + 32 + (kThreads / 32) * GT + (size_t)EC * (rp + 1) + (size_t)EC * (Bg + 1));
}

}  // namespace

extern "C" int tvts_contrastive_sortce_fused_supported(int64_t Bg, int64_t E) { return Bg >= 1 && Bg <= kMaxBg && E >= 1 && E <= 2 * kThreads; }

extern "C" int tvts_contrastive_sortce_fused(const float* video_all, const float* text_all, int64_t Bg, int64_t E, int64_t row0, int64_t nloc,
                                             float temperature, float eps, const float* logits, const int64_t* labels, int64_t R, int64_t C,
                                             float ce_weight, float* loss1, float* loss2, float* d_video, float* d_text, float* dlogits,
                                             void* stream) {
  TVTS_REQUIRE(tvts_contrastive_sortce_fused_supported(Bg, E), "contrastive_sortce_fused: Bg=%lld (<= %d) / E=%lld (<= %d) unsupported",
               (long long)Bg, kMaxBg, (long long)E, 2 * kThreads);
  TVTS_REQUIRE(video_all && text_all && loss1 && d_video && d_text && temperature > 0.f, "contrastive_sortce_fused: bad arguments");
  TVTS_REQUIRE(row0 >= 0 && nloc >= 0 && row0 + nloc <= Bg, "contrastive_sortce_fused: local row range out of bounds");
  TVTS_REQUIRE(logits == nullptr || (labels && loss2 && R > 0 && C > 0), "contrastive_sortce_fused: sort-CE arguments");
  int csize = Bg >= 32 ? 8 : (Bg >= 16 ? 4 : 1);
  const int rp = (int)((Bg + csize - 1) / csize);
  TVTS_REQUIRE(rp <= 32, "contrastive_sortce_fused: rows per CTA");
  FusedArgs p{video_all, text_all, (int)Bg, (int)E, (int)row0, (int)nloc, rp, 1.0f / temperature, eps,
              logits, reinterpret_cast<const long long*>(labels), (int)R, (int)C, ce_weight, loss1, loss2, d_video, d_text, dlogits};
  const size_t smem = fused_smem_bytes((int)Bg, rp);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(contrastive_sortce_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)csize);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, contrastive_sortce_fused_kernel, p);
  tvts_count_launch(1);
  if (le != cudaSuccess) return tvts_set_error(TVTS_ERR_CUDA, "contrastive_sortce_fused launch failed: %s", cudaGetErrorString(le));
  return TVTS_OK;
}
