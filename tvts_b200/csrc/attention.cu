// Grouped multi-head attention for the TVTS hot path, head dim 64, forward + backward:
//   mode FULL  : every token attends to every token (sort head, v2/model/sort_transformer.py:9-13,43-57), optionally
//                causal (CLIP text tower, v2/CLIP/clip/model.py:185-188 with the mask of :330-336)
//   mode SPACE : divided space attention of VarAttention (v2/model/video_encoder_ViT_B_16.py:38-76, '(b f) n d'):
//                a patch token attends to [CLS ; the n kept tokens of its own frame]; CLS attends to all N tokens
//   mode TIME  : divided time attention ('(b n) f d'): a patch token attends to [CLS ; the T tokens of its slot]
// The reference materialises q_/k_/v_/cls_k/cls_v copies with rearrange/repeat/cat; here the grouping is pure index
// arithmetic on the packed qkv buffer [B, N, 3, H, 64] written by the qkv GEMM, so nothing is copied.
//
// One kernel shape serves forward, dQ and dK/dV ("vector-stationary"): a CTA owns up to 32/64 STATIONARY tokens of one
// (batch, head, group) and streams the group's other side through shared memory in tiles of 64 rows:
//   forward / dQ : stationary = queries, streamed = keys (+values)
//   dK,dV        : stationary = keys,    streamed = queries (+dO)       (same index sets: the relation is symmetric)
// dots are computed lane-per-streamed-row (conflict-free padded smem rows), accumulations lane-per-2-dims.
// fp32 math, bf16 I/O, exact softmax (online max/sum in forward; saved log-sum-exp in backward).
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

constexpr int HD = 64;        // head dim
constexpr int KT = 64;        // streamed rows per smem tile
constexpr int PITCH = 33;     // u32 per smem row (32 + 1 pad)
constexpr int kWarps = 4;
constexpr int SC_FWD = 64;    // stationary rows per CTA (forward)
constexpr int SC_BWD = 32;    // stationary rows per CTA (backward)

struct AttnShape {
  int B, N, H;
  int mode;    // 0 full, 1 space, 2 time
  int T, n;    // frames, kept tokens per frame (modes 1, 2)
  int causal;  // mode 0 only
  float scale;
};

struct Sets {
  int st_base, st_stride, st_count;
  int sm_has0, sm_base, sm_stride, sm_count;
};

__host__ __device__ inline int chunks_per_group(const AttnShape& a, int sc) {
  if (a.mode == 0) return (a.N + sc - 1) / sc;
  const int len = a.mode == 1 ? a.n : a.T;
  return (len + sc - 1) / sc;
}
__host__ __device__ inline int num_blocks_x(const AttnShape& a, int sc) {
  if (a.mode == 0) return chunks_per_group(a, sc);
  const int groups = a.mode == 1 ? a.T : a.n;
  return groups * chunks_per_group(a, sc) + 1;  // + the CLS group
}

__device__ inline Sets decode_sets(const AttnShape& a, int bx, int sc) {
  Sets s;
  if (a.mode == 0) {
    s.st_base = bx * sc; s.st_stride = 1; s.st_count = min(sc, a.N - s.st_base);
    s.sm_has0 = 0; s.sm_base = 0; s.sm_stride = 1; s.sm_count = a.N;
    return s;
  }
  const int cpg = chunks_per_group(a, sc);
  const int groups = a.mode == 1 ? a.T : a.n;
  const int g = bx / cpg, c = bx - g * cpg;
  if (g >= groups) {  // CLS token <-> all tokens
    s.st_base = 0; s.st_stride = 1; s.st_count = 1;
    s.sm_has0 = 0; s.sm_base = 0; s.sm_stride = 1; s.sm_count = a.N;
    return s;
  }
  if (a.mode == 1) {
    s.st_base = 1 + g * a.n + c * sc; s.st_stride = 1; s.st_count = min(sc, a.n - c * sc);
    s.sm_has0 = 1; s.sm_base = 1 + g * a.n; s.sm_stride = 1; s.sm_count = a.n;
  } else {
    s.st_base = 1 + g + c * sc * a.n; s.st_stride = a.n; s.st_count = min(sc, a.T - c * sc);
    s.sm_has0 = 1; s.sm_base = 1 + g; s.sm_stride = a.n; s.sm_count = a.T;
  }
  return s;
}

__device__ __forceinline__ int streamed_token(const Sets& s, int k) {
  return (s.sm_has0 && k == 0) ? 0 : s.sm_base + (k - s.sm_has0) * s.sm_stride;
}

// load 64 bf16 (one head row, 128 B) as 64 floats into registers; all lanes read the same address (broadcast)
__device__ __forceinline__ void load_row_f32(const bf16* p, float* r, float mul) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u = reinterpret_cast<const uint4*>(p)[i];
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r[8 * i + 0] = a.x * mul; r[8 * i + 1] = a.y * mul; r[8 * i + 2] = b.x * mul; r[8 * i + 3] = b.y * mul;
    r[8 * i + 4] = c.x * mul; r[8 * i + 5] = c.y * mul; r[8 * i + 6] = d.x * mul; r[8 * i + 7] = d.y * mul;
  }
}

// cooperative tile load: rows [k0, k0+cnt) of the streamed list, 64 bf16 each, into padded smem (u32 pairs)
__device__ __forceinline__ void load_tile(uint32_t* dst, const bf16* base, long long row_stride, const Sets& s, int k0, int cnt) {
  for (int i = threadIdx.x; i < cnt * 8; i += kWarps * 32) {
    const int r = i >> 3, c = i & 7;
    const int tok = streamed_token(s, k0 + r);
    uint4 u = reinterpret_cast<const uint4*>(base + (long long)tok * row_stride)[c];
    uint32_t* d = dst + r * PITCH + c * 4;
    d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
  }
}

__device__ __forceinline__ float dot_row(const float* a, const uint32_t* row) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    float2 k = unpack_bf16x2(row[w]);
    s0 = fmaf(a[2 * w], k.x, s0);
    s1 = fmaf(a[2 * w + 1], k.y, s1);
  }
  return s0 + s1;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kWarps * 32) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                               float* __restrict__ lse, AttnShape a) {
  __shared__ uint32_t Ks[KT * PITCH];
  __shared__ uint32_t Vs[KT * PITCH];
  __shared__ float st_acc[SC_FWD][HD];
  __shared__ float st_m[SC_FWD], st_l[SC_FWD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const Sets s = decode_sets(a, blockIdx.x, SC_FWD);
  const long long rs = 3LL * a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + (long long)a.H * HD;
  const bf16* vb = kb + (long long)a.H * HD;
  const int total = s.sm_has0 + s.sm_count;
  int t_end = (total + KT - 1) / KT;
  if (a.causal) {
    const int last = s.st_base + (s.st_count - 1) * s.st_stride;  // largest query index of the chunk
    t_end = min(t_end, last / KT + 1);
  }
  for (int t = 0; t < t_end; ++t) {
    const int k0 = t * KT;
    const int cnt = min(KT, total - k0);
    __syncthreads();
    load_tile(Ks, kb, rs, s, k0, cnt);
    load_tile(Vs, vb, rs, s, k0, cnt);
    __syncthreads();
    for (int r = warp; r < s.st_count; r += kWarps) {
      const int qi = s.st_base + r * s.st_stride;
      float q[HD];
      load_row_f32(qb + (long long)qi * rs, q, a.scale);
      float m_old = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
      if (t > 0) { m_old = st_m[r]; l = st_l[r]; acc0 = st_acc[r][2 * lane]; acc1 = st_acc[r][2 * lane + 1]; }
      float sc[KT / 32];
      float mt = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < KT / 32; ++kk) {
        const int key = kk * 32 + lane;
        float v = -INFINITY;
        if (key < cnt) {
          const int tok = streamed_token(s, k0 + key);
          if (!a.causal || tok <= qi) v = dot_row(q, Ks + key * PITCH);
        }
        sc[kk] = v;
        mt = fmaxf(mt, v);
      }
      mt = warp_max(mt);
      const float m_new = fmaxf(m_old, mt);
      float corr = 0.f, psum = 0.f;
      if (m_new != -INFINITY) {
        corr = __expf(m_old - m_new);
#pragma unroll
        for (int kk = 0; kk < KT / 32; ++kk) {
          sc[kk] = __expf(sc[kk] - m_new);
          psum += sc[kk];
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < KT / 32; ++kk) sc[kk] = 0.f;
      }
      psum = warp_sum(psum);
      l = l * corr + psum;
      acc0 *= corr; acc1 *= corr;
#pragma unroll
      for (int kk = 0; kk < KT / 32; ++kk) {
        const int lim = min(32, cnt - kk * 32);
        for (int j = 0; j < lim; ++j) {
          const float p = __shfl_sync(0xffffffffu, sc[kk], j);
          const float2 v = unpack_bf16x2(Vs[(kk * 32 + j) * PITCH + lane]);
          acc0 = fmaf(p, v.x, acc0);
          acc1 = fmaf(p, v.y, acc1);
        }
      }
      if (t + 1 < t_end) {
        st_acc[r][2 * lane] = acc0; st_acc[r][2 * lane + 1] = acc1;
        if (lane == 0) { st_m[r] = m_new; st_l[r] = l; }
      } else {
        const float inv = 1.0f / l;
        reinterpret_cast<uint32_t*>(out + ((long long)b * a.N + qi) * a.H * HD + (long long)h * HD)[lane] =
            pack_bf16x2(acc0 * inv, acc1 * inv);
        if (lane == 0) lse[((long long)b * a.H + h) * a.N + qi] = m_new + __logf(l);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO * O)
__global__ void attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta, int B, int N,
                                  int H) {
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * N * H) return;
  const int h = (int)(w % H);
  const long long bn = w / H;
  const int i = (int)(bn % N);
  const int b = (int)(bn / N);
  const float2 o = unpack_bf16x2(reinterpret_cast<const uint32_t*>(out + w * HD)[lane]);
  const float2 d = unpack_bf16x2(reinterpret_cast<const uint32_t*>(dout + w * HD)[lane]);
  const float s = warp_sum(o.x * d.x + o.y * d.y);
  if (lane == 0) delta[((long long)b * H + h) * N + i] = s;
}

// ------------------------------------------------------------------------------------------------ backward
// ROLE 0: stationary = query i  (a = q_i, b = dO_i), streamed rows X = K_j, Y = V_j      -> dq_i = scale * sum_j ds_ij K_j
// ROLE 1: stationary = key j    (a = k_j, b = v_j),  streamed rows X = Q_i, Y = dO_i     -> dk_j = scale * sum_i ds_ij Q_i
//                                                                                            dv_j = sum_i p_ij dO_i
// with p_ij = exp(scale q_i.k_j - lse_i), ds_ij = p_ij (dO_i.v_j - delta_i).
template <int ROLE>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               bf16* __restrict__ dqkv, AttnShape a) {
  __shared__ uint32_t Xs[KT * PITCH];
  __shared__ uint32_t Ys[KT * PITCH];
  __shared__ float Ls[KT], Ds[KT];
  __shared__ float st_acc[SC_BWD][ROLE == 1 ? 2 * HD : HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const Sets s = decode_sets(a, blockIdx.x, SC_BWD);
  const long long rs = 3LL * a.H * HD;   // qkv token stride
  const long long ro = (long long)a.H * HD;  // out/dout token stride
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + ro;
  const bf16* vb = kb + ro;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  bf16* dq_b = dqkv + (long long)b * a.N * rs + (long long)h * HD;

  const bf16* xs_base = ROLE == 0 ? kb : qb;
  const bf16* ys_base = ROLE == 0 ? vb : dob;
  const long long ys_stride = ROLE == 0 ? rs : ro;

  const int total = s.sm_has0 + s.sm_count;
  int t_begin = 0, t_end = (total + KT - 1) / KT;
  if (a.causal) {
    if (ROLE == 0) t_end = min(t_end, (s.st_base + (s.st_count - 1) * s.st_stride) / KT + 1);
    else t_begin = s.st_base / KT;  // queries before the first key of the chunk never see it
  }
  for (int t = t_begin; t < t_end; ++t) {
    const int k0 = t * KT;
    const int cnt = min(KT, total - k0);
    __syncthreads();
    load_tile(Xs, xs_base, rs, s, k0, cnt);
    load_tile(Ys, ys_base, ys_stride, s, k0, cnt);
    if (ROLE == 1) {
      for (int i = threadIdx.x; i < cnt; i += kWarps * 32) {
        const int tok = streamed_token(s, k0 + i);
        Ls[i] = lse_b[tok];
        Ds[i] = delta_b[tok];
      }
    }
    __syncthreads();
    for (int r = warp; r < s.st_count; r += kWarps) {
      const int si = s.st_base + r * s.st_stride;
      float av[HD], bv[HD];
      if (ROLE == 0) {
        load_row_f32(qb + (long long)si * rs, av, 1.0f);
        load_row_f32(dob + (long long)si * ro, bv, 1.0f);
      } else {
        load_row_f32(kb + (long long)si * rs, av, 1.0f);
        load_row_f32(vb + (long long)si * rs, bv, 1.0f);
      }
      float lse_s = 0.f, delta_s = 0.f;
      if (ROLE == 0) { lse_s = lse_b[si]; delta_s = delta_b[si]; }
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      if (t > t_begin) {
        acc0 = st_acc[r][2 * lane]; acc1 = st_acc[r][2 * lane + 1];
        if (ROLE == 1) { acc2 = st_acc[r][HD + 2 * lane]; acc3 = st_acc[r][HD + 2 * lane + 1]; }
      }
      float pv[KT / 32], dsv[KT / 32];
#pragma unroll
      for (int kk = 0; kk < KT / 32; ++kk) {
        const int row = kk * 32 + lane;
        float p = 0.f, ds = 0.f;
        if (row < cnt) {
          const int tok = streamed_token(s, k0 + row);
          const bool ok = !a.causal || (ROLE == 0 ? tok <= si : tok >= si);
          if (ok) {
            const float sdot = dot_row(av, Xs + row * PITCH) * a.scale;
            const float dp = dot_row(bv, Ys + row * PITCH);
            const float l_ = ROLE == 0 ? lse_s : Ls[row];
            const float d_ = ROLE == 0 ? delta_s : Ds[row];
            p = __expf(sdot - l_);
            ds = p * (dp - d_);
          }
        }
        pv[kk] = p;
        dsv[kk] = ds;
      }
#pragma unroll
      for (int kk = 0; kk < KT / 32; ++kk) {
        const int lim = min(32, cnt - kk * 32);
        for (int j = 0; j < lim; ++j) {
          const float ds = __shfl_sync(0xffffffffu, dsv[kk], j);
          const float2 x = unpack_bf16x2(Xs[(kk * 32 + j) * PITCH + lane]);
          acc0 = fmaf(ds, x.x, acc0);
          acc1 = fmaf(ds, x.y, acc1);
          if (ROLE == 1) {
            const float p = __shfl_sync(0xffffffffu, pv[kk], j);
            const float2 y = unpack_bf16x2(Ys[(kk * 32 + j) * PITCH + lane]);
            acc2 = fmaf(p, y.x, acc2);
            acc3 = fmaf(p, y.y, acc3);
          }
        }
      }
      if (t + 1 < t_end) {
        st_acc[r][2 * lane] = acc0; st_acc[r][2 * lane + 1] = acc1;
        if (ROLE == 1) { st_acc[r][HD + 2 * lane] = acc2; st_acc[r][HD + 2 * lane + 1] = acc3; }
      } else {
        bf16* o = dq_b + (long long)si * rs + (ROLE == 0 ? 0 : ro);
        reinterpret_cast<uint32_t*>(o)[lane] = pack_bf16x2(acc0 * a.scale, acc1 * a.scale);
        if (ROLE == 1) reinterpret_cast<uint32_t*>(o + ro)[lane] = pack_bf16x2(acc2, acc3);
      }
    }
  }
}

int check_shape(const AttnShape& a, int64_t d) {
  TVTS_REQUIRE(d == HD, "attention: head dim %lld unsupported (only 64)", (long long)d);
  TVTS_REQUIRE(a.B > 0 && a.N > 0 && a.H > 0, "attention: empty shape");
  TVTS_REQUIRE(a.mode >= 0 && a.mode <= 2, "attention: bad mode %d", a.mode);
  if (a.mode != 0) {
    TVTS_REQUIRE(a.T > 0 && a.n > 0 && a.N == 1 + a.T * a.n, "attention: N=%d != 1 + T*n (T=%d n=%d)", a.N, a.T, a.n);
    TVTS_REQUIRE(!a.causal, "attention: causal only valid in full mode");
  }
  TVTS_REQUIRE(a.H <= 65535 && a.B <= 65535, "attention: grid limits");
  return TVTS_OK;
}

}  // namespace

extern "C" int tvts_attn_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T,
                             int64_t n, int64_t causal, float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale};
  if (B == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && lse, "attn_fwd: null pointer");
  dim3 grid(num_blocks_x(a, SC_FWD), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<<<grid, kWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)qkv, (bf16*)out, lse, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int64_t B,
                             int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale,
                             void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale};
  if (B == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv, "attn_bwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)B * N * H;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta_ws, (int)B, (int)N, (int)H);
  TVTS_LAUNCH_CHECK();
  dim3 grid(num_blocks_x(a, SC_BWD), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<0><<<grid, kWarps * 32, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  attn_bwd_kernel<1><<<grid, kWarps * 32, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
