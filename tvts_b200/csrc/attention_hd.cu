// Streamed grouped attention for head dims other than 64 (TVTSv2 ViT-H/14: D 1280 / 16 heads = 80,
// v2/model/model_dist_TVTSv2_ViT_H_14.py:43-45 + v2/OpenCLIP/model_configs/ViT-H-14.json), forward + backward.
//
// Same algorithm, index sets (attention_common.cuh: FULL / SPACE / TIME with the CLS row and column) and launch geometry as the
// streamed kernels of attention.cu, templated on the head dim HD (a multiple of 16, <= 128):
//   * a token's head slice is HD/8 16-byte chunks; shared-memory tiles keep 64 rows at a pitch of 128 B (HD <= 64) or 256 B, with the
//     16-byte chunk index XOR-ed with (row & 7) so that ldmatrix stays bank-conflict free; chunks past HD/8 are never touched
//   * S = Q K^T runs HD/16 k-steps, O = P V / dQ / dK / dV produce HD/8 8-wide n-tiles per warp
// tvts_attn_fwd / tvts_attn_bwd (attention.cu) forward here whenever d != 64; the HD = 64 instantiation exists only so that the
// generic code can be checked against the specialised kernels (tests call tvts_attn_generic_* with d = 64 directly).
#include "attention_common.cuh"
#ifndef TVTS_HOST_SHIM
#include "../../include/tvts_b200.h"
#endif

namespace {

template <int HD>
struct Geo {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 128, "head dim must be a multiple of 16, at most 128");
  static constexpr int CH = HD / 8;                    // 16-byte chunks per row
  static constexpr int KS = HD / 16;                   // k-steps of a product that contracts over the head dim
  static constexpr int ND = HD / 8;                    // 8-wide n-tiles of a product whose columns are the head dim
  static constexpr int PITCH = HD <= 64 ? 128 : 256;   // shared-memory row pitch in bytes
  static constexpr int TILE = 64 * PITCH;              // one 64-row tile
};

// byte offset of 16-byte chunk `c` of row `r` in a tile of pitch P (c ^ (r & 7) stays inside c's group of 8 chunks = 128 B)
template <int P>
__device__ __forceinline__ uint32_t swzp(int r, int c) { return (uint32_t)(r * P + ((c ^ (r & 7)) << 4)); }

// A fragments: 16 rows starting at row0, all HD dims
template <int HD>
__device__ __forceinline__ void g_load_a(uint32_t tile, int row0, int lane, uint32_t (&a)[Geo<HD>::KS][4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < Geo<HD>::KS; ++kk)
    ldsm_x4(tile + swzp<Geo<HD>::PITCH>(r, 2 * kk + (lane >> 4)), a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}
// B fragments, tile ROWS as the n index (S = A . tile^T): n-tiles j, j+1 (rows 8j .. 8j+15), k-step kk (dims 16kk ..)
template <int HD>
__device__ __forceinline__ void g_load_b_rows(uint32_t tile, int j, int kk, int lane, uint32_t& b0, uint32_t& b1, uint32_t& c0, uint32_t& c1) {
  const int r = 8 * j + (lane & 7) + ((lane >> 4) & 1) * 8;
  ldsm_x4(tile + swzp<Geo<HD>::PITCH>(r, 2 * kk + ((lane >> 3) & 1)), b0, b1, c0, c1);
}
// B fragments, tile ROWS as the k index (O = P . tile): k-step kk (rows 16kk .. 16kk+15), n-tiles jd, jd+1 (dims 8jd ..)
template <int HD>
__device__ __forceinline__ void g_load_b_cols(uint32_t tile, int kk, int jd, int lane, uint32_t& b0, uint32_t& b1, uint32_t& c0, uint32_t& c1) {
  const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
  ldsm_x4_t(tile + swzp<Geo<HD>::PITCH>(r, jd + (lane >> 4)), b0, b1, c0, c1);
}
// cooperative 64-row tile load (rows >= cnt are zero-filled); token(r) gives the token index of tile row r
template <int HD, typename TokFn>
__device__ __forceinline__ void g_load_tile(uint32_t tile, const bf16* base, long long row_stride, int cnt, TokFn token) {
  constexpr int CH = Geo<HD>::CH;
  for (int i = threadIdx.x; i < 64 * CH; i += kThreads) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = r < cnt;
    const long long tok = ok ? token(r) : 0;
    cp_async16(tile + swzp<Geo<HD>::PITCH>(r, c), base + tok * row_stride + c * 8, ok);
  }
}

// ================================================================================================ forward
template <int HD>
__global__ void __launch_bounds__(kThreads) attn_hd_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                                               AttnShape a, const int* __restrict__ klen) {
  using G = Geo<HD>;
  TVTS_DYN_SMEM(uint8_t, smem, 128);   // Q | K0 | V0 | K1 | V1
  const uint32_t sQ = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  Sets s = decode_sets(a, blockIdx.x);
  if (klen) s.sm_count = min(s.sm_count, klen[b]);   // key padding (mode FULL): keys >= klen[b] are simply never streamed
  const long long rs = 3LL * a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + (long long)a.H * HD;
  const bf16* vb = kb + (long long)a.H * HD;
  const int total = s.sm_has0 + s.sm_count;
  int t_end = (total + BN - 1) / BN;
  if (a.causal) t_end = min(t_end, (s.st_base + (s.st_count - 1) * s.st_stride) / BN + 1);

  g_load_tile<HD>(sQ, qb, rs, s.st_count, [&](int r) { return s.st_base + r * s.st_stride; });
  auto issue = [&](int t) {
    const uint32_t sK = sQ + (1 + 2 * (t & 1)) * G::TILE, sV = sK + G::TILE;
    const int k0 = t * BN, cnt = min(BN, total - k0);
    g_load_tile<HD>(sK, kb, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    g_load_tile<HD>(sV, vb, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
  };
  issue(0);
  cp_async_commit();

  const bool warp_active = warp * 16 < s.st_count;   // warp-uniform
  uint32_t qa[G::KS][4];
  float o[G::ND][4];
#pragma unroll
  for (int j = 0; j < G::ND; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl2 = a.scale * LOG2E;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int qtok0 = s.st_base + row0 * s.st_stride, qtok1 = s.st_base + row1 * s.st_stride;

  for (int t = 0; t < t_end; ++t) {
    if (t + 1 < t_end) issue(t + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) g_load_a<HD>(sQ, warp * 16, lane, qa);
    if (warp_active) {
      const uint32_t sK = sQ + (1 + 2 * (t & 1)) * G::TILE, sV = sK + G::TILE;
      const int k0 = t * BN, cnt = min(BN, total - k0);
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        if (8 * j < cnt) {     // CTA-uniform: key tiles past the end of a short group are never computed
#pragma unroll
          for (int kk = 0; kk < G::KS; ++kk) {
            uint32_t b0, b1, c0, c1;
            g_load_b_rows<HD>(sK, j, kk, lane, b0, b1, c0, c1);
            mma16816(sc[j], qa[kk], b0, b1);
            mma16816(sc[j + 1], qa[kk], c0, c1);
          }
        }
      }
      // mask + row max
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (8 * j >= cnt) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY; continue; }   // CTA-uniform
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = 8 * j + 2 * t4 + e;
          bool ok = key < cnt;
          bool ok0 = ok, ok1 = ok;
          if (a.causal && ok) {
            const int tok = streamed_token(s, k0 + key);
            ok0 = tok <= qtok0; ok1 = tok <= qtok1;
          }
          if (!ok0) sc[j][e] = -INFINITY;
          if (!ok1) sc[j][2 + e] = -INFINITY;
          mx0 = fmaxf(mx0, sc[j][e]);
          mx1 = fmaxf(mx1, sc[j][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float mu0 = mn0 == -INFINITY ? 0.f : mn0, mu1 = mn1 == -INFINITY ? 0.f : mn1;
      const float corr0 = exp2f((m0 - mu0) * sl2), corr1 = exp2f((m1 - mu1) * sl2);   // m = -inf -> 0
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (8 * j >= cnt) { pa[j >> 1][(j & 1) * 2] = 0u; pa[j >> 1][(j & 1) * 2 + 1] = 0u; continue; }
        const float p00 = exp2f((sc[j][0] - mu0) * sl2), p01 = exp2f((sc[j][1] - mu0) * sl2);
        const float p10 = exp2f((sc[j][2] - mu1) * sl2), p11 = exp2f((sc[j][3] - mu1) * sl2);
        ps0 += p00 + p01; ps1 += p10 + p11;
        pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p00, p01);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p10, p11);
      }
      l0 = l0 * corr0 + ps0; l1 = l1 * corr1 + ps1;   // per-thread partial sums; reduced across the quad at the end
#pragma unroll
      for (int j = 0; j < G::ND; ++j) { o[j][0] *= corr0; o[j][1] *= corr0; o[j][2] *= corr1; o[j][3] *= corr1; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (16 * kk < cnt) {
#pragma unroll
          for (int jd = 0; jd < G::ND; jd += 2) {
            uint32_t b0, b1, c0, c1;
            g_load_b_cols<HD>(sV, kk, jd, lane, b0, b1, c0, c1);
            mma16816(o[jd], pa[kk], b0, b1);
            mma16816(o[jd + 1], pa[kk], c0, c1);
          }
        }
      }
    }
    __syncthreads();   // everyone done with this stage's K/V before it is refilled
  }
  cp_async_wait<0>();
  if (!warp_active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  // stage the warp's 16 output rows in its slice of the Q tile (its fragments are in registers), then 16-byte row stores
  __syncwarp();
#pragma unroll
  for (int j = 0; j < G::ND; ++j) {
    const uint32_t w0 = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0), w1 = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    *reinterpret_cast<uint32_t*>(smem + swzp<G::PITCH>(row0, j) + 4 * t4) = w0;
    *reinterpret_cast<uint32_t*>(smem + swzp<G::PITCH>(row1, j) + 4 * t4) = w1;
  }
  __syncwarp();
  const long long ro = (long long)a.H * HD;
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
  for (int i = lane; i < 16 * G::CH; i += 32) {
    const int rr = i / G::CH, c = i - rr * G::CH;
    const int r = warp * 16 + rr;
    if (r < s.st_count) {
      const uint4 v = *reinterpret_cast<const uint4*>(smem + swzp<G::PITCH>(r, c));
      *reinterpret_cast<uint4*>(ob + (long long)(s.st_base + r * s.st_stride) * ro + c * 8) = v;
    }
  }
  if (t4 == 0) {
    float* lb = lse + ((long long)b * a.H + h) * a.N;
    if (row0 < s.st_count) lb[qtok0] = m0 * a.scale + __logf(l0);
    if (row1 < s.st_count) lb[qtok1] = m1 * a.scale + __logf(l1);
  }
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO * O)
template <int HD>
__global__ void __launch_bounds__(256) attn_hd_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                            float* __restrict__ delta, int B, int N, int H) {
  // 8 lanes per (token, head) row of HD bf16: 16-byte loads of O and dO, chunks strided over the octet, reduce inside the octet
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool ok = w < (long long)B * N * H;
  float s = 0.f;
  if (ok) {
    for (int c = sub; c < Geo<HD>::CH; c += 8) {
      const uint4 o = reinterpret_cast<const uint4*>(out + w * HD)[c];
      const uint4 d = reinterpret_cast<const uint4*>(dout + w * HD)[c];
      const uint32_t ov[4] = {o.x, o.y, o.z, o.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(ov[i]), y = unpack_bf16x2(dv[i]);
        s = fmaf(x.x, y.x, s);
        s = fmaf(x.y, y.y, s);
      }
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) {
    const int h = (int)(w % H);
    const long long bn = w / H;
    const int i = (int)(bn % N);
    const int b = (int)(bn / N);
    delta[((long long)b * H + h) * N + i] = s;
  }
}

// ================================================================================================ backward
// ROLE 0: stationary = query i (tiles Q, dO), streamed X = K, Y = V        -> dq_i = scale * sum_j ds_ij k_j
// ROLE 1: stationary = key j   (tiles K, V),  streamed X = Q, Y = dO       -> dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i p_ij do_i
// with p_ij = exp(scale q_i.k_j - lse_i), ds_ij = p_ij (do_i.v_j - delta_i)   (see attn_bwd_kernel in attention.cu)
template <int HD, int ROLE>
__global__ void __launch_bounds__(kThreads) attn_hd_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               bf16* __restrict__ dqkv, AttnShape a, const int* __restrict__ klen) {
  using G = Geo<HD>;
  TVTS_DYN_SMEM(uint8_t, smem, 128);   // 4 tiles + lse/delta of the streamed rows for 2 stages
  // [0] stationary tile 0 (Q or K) -> after the fragment load reused as X stage 1
  // [1] stationary tile 1 (dO or V) -> reused as Y stage 1
  // [2] X stage 0, [3] Y stage 0
  const uint32_t s0 = smem_u32(smem);
  float* LsDs = reinterpret_cast<float*>(smem + 4 * G::TILE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  Sets s = decode_sets(a, blockIdx.x, ROLE == 1);
  // key padding (mode FULL): ROLE 0 never streams keys >= klen[b]; ROLE 1 masks its stationary keys >= klen[b] (their dk / dv are 0)
  const int kl = klen ? klen[b] : 0x7fffffff;
  if (ROLE == 0 && klen) s.sm_count = min(s.sm_count, kl);
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
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
  int t_begin = 0, t_end = (total + BN - 1) / BN;
  if (a.causal) {
    if (ROLE == 0) t_end = min(t_end, (s.st_base + (s.st_count - 1) * s.st_stride) / BN + 1);
    else t_begin = s.st_base / BN;  // queries before the first key of the chunk never see it
  }
  auto st_tok = [&](int r) { return s.st_base + r * s.st_stride; };
  auto issue = [&](int t, int stage) {
    const uint32_t sX = s0 + (stage == 0 ? 2 : 0) * G::TILE, sY = sX + G::TILE;
    const int k0 = t * BN, cnt = min(BN, total - k0);
    g_load_tile<HD>(sX, xs_base, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    g_load_tile<HD>(sY, ys_base, ys_stride, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    if (ROLE == 1) {
      for (int i = threadIdx.x; i < BN; i += kThreads) {
        float l_ = 0.f, d_ = 0.f;
        if (i < cnt) { const int tok = streamed_token(s, k0 + i); l_ = lse_b[tok]; d_ = delta_b[tok]; }
        LsDs[stage * 2 * BN + i] = l_ * LOG2E;
        LsDs[stage * 2 * BN + BN + i] = d_;
      }
    }
  };
  if (ROLE == 0) {
    g_load_tile<HD>(s0, qb, rs, s.st_count, st_tok);
    g_load_tile<HD>(s0 + G::TILE, dob, ro, s.st_count, st_tok);
  } else {
    g_load_tile<HD>(s0, kb, rs, s.st_count, st_tok);
    g_load_tile<HD>(s0 + G::TILE, vb, rs, s.st_count, st_tok);
  }
  issue(t_begin, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t fa[G::KS][4], fb[G::KS][4];   // ROLE 0: Q, dO ; ROLE 1: K, V   (A operands, 16 rows of this warp)
  g_load_a<HD>(s0, warp * 16, lane, fa);
  g_load_a<HD>(s0 + G::TILE, warp * 16, lane, fb);
  __syncthreads();                       // tiles 0,1 may now be overwritten (stage 1)

  const bool warp_active = warp * 16 < s.st_count;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int stok0 = st_tok(row0), stok1 = st_tok(row1);
  float lse0 = 0.f, lse1 = 0.f, dl0 = 0.f, dl1 = 0.f;
  if (ROLE == 0) {
    if (row0 < s.st_count) { lse0 = lse_b[stok0] * LOG2E; dl0 = delta_b[stok0]; }
    if (row1 < s.st_count) { lse1 = lse_b[stok1] * LOG2E; dl1 = delta_b[stok1]; }
  }
  const float sl2 = a.scale * LOG2E;
  float acc0[G::ND][4], acc1[ROLE == 1 ? G::ND : 1][4];   // ROLE 0: acc0 = dQ ; ROLE 1: acc0 = dK, acc1 = dV
#pragma unroll
  for (int j = 0; j < G::ND; ++j) acc0[j][0] = acc0[j][1] = acc0[j][2] = acc0[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < (ROLE == 1 ? G::ND : 1); ++j) acc1[j][0] = acc1[j][1] = acc1[j][2] = acc1[j][3] = 0.f;

  for (int t = t_begin; t < t_end; ++t) {
    const int stage = (t - t_begin) & 1;
    if (t + 1 < t_end) issue(t + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (warp_active) {
      const uint32_t sX = s0 + (stage == 0 ? 2 : 0) * G::TILE, sY = sX + G::TILE;
      const float* Ls = LsDs + stage * 2 * BN;
      const float* Ds = Ls + BN;
      const int k0 = t * BN, cnt = min(BN, total - k0);
#pragma unroll
      for (int half = 0; half < 2; ++half) {   // 32 streamed rows at a time (register pressure)
        if (32 * half >= cnt) break;           // CTA-uniform
        float sc[4][4], dp[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (32 * half + 8 * j < cnt) {
#pragma unroll
            for (int kk = 0; kk < G::KS; ++kk) {
              uint32_t b0, b1, c0, c1;
              g_load_b_rows<HD>(sX, half * 4 + j, kk, lane, b0, b1, c0, c1);
              mma16816(sc[j], fa[kk], b0, b1);
              mma16816(sc[j + 1], fa[kk], c0, c1);
              g_load_b_rows<HD>(sY, half * 4 + j, kk, lane, b0, b1, c0, c1);
              mma16816(dp[j], fb[kk], b0, b1);
              mma16816(dp[j + 1], fb[kk], c0, c1);
            }
          }
        }
        uint32_t pa[2][4], dsa[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (32 * half + 8 * j >= cnt) {                        // CTA-uniform: nothing streamed there
            dsa[j >> 1][(j & 1) * 2] = 0u; dsa[j >> 1][(j & 1) * 2 + 1] = 0u;
            pa[j >> 1][(j & 1) * 2] = 0u; pa[j >> 1][(j & 1) * 2 + 1] = 0u;
            continue;
          }
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 32 * half + 8 * j + 2 * t4 + e;      // streamed row index inside the tile
            bool ok = col < cnt, ok0 = ok, ok1 = ok;
            if (a.causal && ok) {
              const int tok = streamed_token(s, k0 + col);
              if (ROLE == 0) { ok0 = tok <= stok0; ok1 = tok <= stok1; }
              else { ok0 = tok >= stok0; ok1 = tok >= stok1; }
            }
            if (ROLE == 1) { ok0 = ok0 && stok0 < kl; ok1 = ok1 && stok1 < kl; }
            const float la = ROLE == 0 ? lse0 : Ls[col], lb = ROLE == 0 ? lse1 : Ls[col];
            const float da = ROLE == 0 ? dl0 : Ds[col], db = ROLE == 0 ? dl1 : Ds[col];
            p[e] = ok0 ? exp2f(sc[j][e] * sl2 - la) : 0.f;
            p[2 + e] = ok1 ? exp2f(sc[j][2 + e] * sl2 - lb) : 0.f;
            ds[e] = p[e] * (dp[j][e] - da);
            ds[2 + e] = p[2 + e] * (dp[j][2 + e] - db);
          }
          dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
          dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
          pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p[0], p[1]);
          pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (32 * half + 16 * kk >= cnt) break;
#pragma unroll
          for (int jd = 0; jd < G::ND; jd += 2) {
            uint32_t b0, b1, c0, c1;
            g_load_b_cols<HD>(sX, half * 2 + kk, jd, lane, b0, b1, c0, c1);     // ROLE 0: K (dQ += dS K) ; ROLE 1: Q (dK += dS^T Q)
            mma16816(acc0[jd], dsa[kk], b0, b1);
            mma16816(acc0[jd + 1], dsa[kk], c0, c1);
            if (ROLE == 1) {
              g_load_b_cols<HD>(sY, half * 2 + kk, jd, lane, b0, b1, c0, c1);   // dO (dV += P^T dO)
              mma16816(acc1[jd], pa[kk], b0, b1);
              mma16816(acc1[jd + 1], pa[kk], c0, c1);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();   // all warps are past their last tile: tile 2/3 region is free for output staging
  if (!warp_active) return;
  // stage through this warp's private 16-row slice of tile 2 (and tile 3 for dV), then 16-byte row stores
  uint8_t* stg = smem + 2 * G::TILE;
#pragma unroll
  for (int j = 0; j < G::ND; ++j) {
    *reinterpret_cast<uint32_t*>(stg + swzp<G::PITCH>(row0, j) + 4 * t4) = pack_bf16x2(acc0[j][0] * a.scale, acc0[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(stg + swzp<G::PITCH>(row1, j) + 4 * t4) = pack_bf16x2(acc0[j][2] * a.scale, acc0[j][3] * a.scale);
    if (ROLE == 1) {
      *reinterpret_cast<uint32_t*>(stg + G::TILE + swzp<G::PITCH>(row0, j) + 4 * t4) = pack_bf16x2(acc1[j][0], acc1[j][1]);
      *reinterpret_cast<uint32_t*>(stg + G::TILE + swzp<G::PITCH>(row1, j) + 4 * t4) = pack_bf16x2(acc1[j][2], acc1[j][3]);
    }
  }
  __syncwarp();
  for (int i = lane; i < 16 * G::CH; i += 32) {
    const int rr = i / G::CH, c = i - rr * G::CH;
    const int r = warp * 16 + rr;
    if (r < s.st_count) {
      bf16* dst = dq_b + (long long)st_tok(r) * rs + (ROLE == 0 ? 0 : ro) + c * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stg + swzp<G::PITCH>(r, c));
      if (ROLE == 1) *reinterpret_cast<uint4*>(dst + ro) = *reinterpret_cast<const uint4*>(stg + G::TILE + swzp<G::PITCH>(r, c));
    }
  }
}


// ================================================================================================ group-resident kernels
// One CTA holds a whole (batch, head, group) in shared memory -- the divided SPACE attention of the H/14 tower (n = 76 kept patches + the
// CLS key = 77 rows -> 5 warps) or a short full-attention sequence -- and finishes it in ONE pass: warp w owns query rows 16w..16w+15
// (and, in the transposed pass of the backward, key rows 16w..16w+15).  Same algorithm as attn_group_fwd/bwd_kernel in attention.cu
// (head dim 64), on tiles of pitch Geo<HD>::PITCH.  In space mode the CLS query row / CLS key column are produced by a `cls_only` launch
// of the streamed kernels above.
struct GroupSetsHD {
  int q_base, q_stride, nq;                  // queries owned by the group
  int k_has0, k_base, k_stride, nk_patch;    // keys: [token 0 if k_has0] ; nk_patch tokens
};
__device__ __forceinline__ GroupSetsHD hd_group_sets(const AttnShape& a, int g) {
  GroupSetsHD s;
  if (a.mode == 0) { s.q_base = 0; s.q_stride = 1; s.nq = a.N; s.k_has0 = 0; s.k_base = 0; s.k_stride = 1; s.nk_patch = a.N; }
  else { s.q_base = 1 + g * a.n; s.q_stride = 1; s.nq = a.n; s.k_has0 = 1; s.k_base = 1 + g * a.n; s.k_stride = 1; s.nk_patch = a.n; }
  return s;
}

template <int HD, int NW>
__global__ void __launch_bounds__(NW * 32) attn_hd_group_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                    float* __restrict__ lse, AttnShape a) {
  using G = Geo<HD>;
  constexpr int R = NW * 16, P = G::PITCH, CH = G::CH;
  TVTS_DYN_SMEM(uint8_t, smem, 128);                   // Q | K | V tiles of R rows
  const uint32_t sQ = smem_u32(smem), sK = sQ + R * P, sV = sK + R * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const GroupSetsHD s = hd_group_sets(a, blockIdx.x);
  const int nk = s.k_has0 + s.nk_patch;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  for (int i = threadIdx.x; i < R * CH; i += NW * 32) {
    const int r = i / CH, c = i - r * CH;
    const bool okq = r < s.nq, okk = r < nk;
    const long long qt = okq ? s.q_base + (long long)r * s.q_stride : 0;
    const long long kt = (!okk || (s.k_has0 && r == 0)) ? 0 : s.k_base + (long long)(r - s.k_has0) * s.k_stride;
    cp_async16(sQ + swzp<P>(r, c), qb + qt * rs + c * 8, okq);
    cp_async16(sK + swzp<P>(r, c), qb + kt * rs + ro + c * 8, okk);
    cp_async16(sV + swzp<P>(r, c), qb + kt * rs + 2 * ro + c * 8, okk);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (warp * 16 >= s.nq) return;
  uint32_t qa[G::KS][4];
  g_load_a<HD>(sQ, warp * 16, lane, qa);
  float sc[2 * NW][4];
#pragma unroll
  for (int j = 0; j < 2 * NW; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int qtok0 = s.q_base + row0 * s.q_stride, qtok1 = s.q_base + row1 * s.q_stride;
  const int k_lim = a.causal ? min(nk, warp * 16 + 16) : nk;     // causal: keys beyond this warp's last query are never needed
#pragma unroll
  for (int j = 0; j < 2 * NW; j += 2) {
    if (8 * j < k_lim) {
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t b0, b1, c0, c1;
        g_load_b_rows<HD>(sK, j, kk, lane, b0, b1, c0, c1);
        mma16816(sc[j], qa[kk], b0, b1);
        mma16816(sc[j + 1], qa[kk], c0, c1);
      }
    }
  }
  const float sl2 = a.scale * LOG2E;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 2 * NW; ++j) {
    if (8 * j >= k_lim) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY; continue; }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = 8 * j + 2 * t4 + e;
      bool ok0 = key < nk, ok1 = ok0;
      if (a.causal) { ok0 = ok0 && key <= qtok0; ok1 = ok1 && key <= qtok1; }     // mode 0: token index == row index
      if (!ok0) sc[j][e] = -INFINITY;
      if (!ok1) sc[j][2 + e] = -INFINITY;
      mx0 = fmaxf(mx0, sc[j][e]); mx1 = fmaxf(mx1, sc[j][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  if (mx0 == -INFINITY) mx0 = 0.f;     // padded query rows
  if (mx1 == -INFINITY) mx1 = 0.f;
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[NW][4];
#pragma unroll
  for (int j = 0; j < 2 * NW; ++j) {
    if (8 * j >= k_lim) { pa[j >> 1][(j & 1) * 2] = 0u; pa[j >> 1][(j & 1) * 2 + 1] = 0u; continue; }
    const float p00 = exp2f((sc[j][0] - mx0) * sl2), p01 = exp2f((sc[j][1] - mx0) * sl2);
    const float p10 = exp2f((sc[j][2] - mx1) * sl2), p11 = exp2f((sc[j][3] - mx1) * sl2);
    l0 += p00 + p01; l1 += p10 + p11;
    pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p00, p01);
    pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p10, p11);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float o[G::ND][4];
#pragma unroll
  for (int j = 0; j < G::ND; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NW; ++kk) {
    if (16 * kk < k_lim) {
#pragma unroll
      for (int jd = 0; jd < G::ND; jd += 2) {
        uint32_t b0, b1, c0, c1;
        g_load_b_cols<HD>(sV, kk, jd, lane, b0, b1, c0, c1);
        mma16816(o[jd], pa[kk], b0, b1);
        mma16816(o[jd + 1], pa[kk], c0, c1);
      }
    }
  }
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __syncwarp();     // this warp's Q rows are only read by this warp (fragments already in registers): reuse them for staging
#pragma unroll
  for (int j = 0; j < G::ND; ++j) {
    *reinterpret_cast<uint32_t*>(smem + swzp<P>(row0, j) + 4 * t4) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(smem + swzp<P>(row1, j) + 4 * t4) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
  __syncwarp();
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
  for (int i = lane; i < 16 * CH; i += 32) {
    const int rr = i / CH, c = i - rr * CH;
    const int r = warp * 16 + rr;
    if (r < s.nq)
      *reinterpret_cast<uint4*>(ob + (long long)(s.q_base + r * s.q_stride) * ro + c * 8) = *reinterpret_cast<const uint4*>(smem + swzp<P>(r, c));
  }
  if (t4 == 0) {
    float* lb = lse + ((long long)b * a.H + h) * a.N;
    if (row0 < s.nq) lb[qtok0] = mx0 * a.scale + __logf(l0);
    if (row1 < s.nq) lb[qtok1] = mx1 * a.scale + __logf(l1);
  }
}

template <int HD, int NW>
__global__ void __launch_bounds__(NW * 32) attn_hd_group_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                    const float* __restrict__ lse, const float* __restrict__ delta,
                                                                    bf16* __restrict__ dqkv, AttnShape a) {
  using G = Geo<HD>;
  constexpr int R = NW * 16, P = G::PITCH, CH = G::CH;
  TVTS_DYN_SMEM(uint8_t, gsm, 128);                    // Q | K | V | dO | dQ-staging tiles, then lse*log2e [R], delta [R]
  const uint32_t sQ = smem_u32(gsm), sK = sQ + R * P, sV = sK + R * P, sD = sV + R * P;
  uint8_t* gO = gsm + 4 * R * P;
  float* Ls = reinterpret_cast<float*>(gsm + 5 * R * P);
  float* Ds = Ls + R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const GroupSetsHD s = hd_group_sets(a, blockIdx.x);
  const int nk = s.k_has0 + s.nk_patch;
  const int nq_all = s.nq + s.k_has0;                 // space mode: the CLS query is row nq of the Q / dO tiles
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  bf16* dq_b = dqkv + (long long)b * a.N * rs + (long long)h * HD;
  auto q_token = [&](int r) -> long long { return r < s.nq ? s.q_base + (long long)r * s.q_stride : 0; };   // r == nq: CLS query
  for (int i = threadIdx.x; i < R * CH; i += NW * 32) {
    const int r = i / CH, c = i - r * CH;
    const bool okq = r < nq_all, okk = r < nk;
    const long long qt = okq ? q_token(r) : 0;
    const long long kt = (!okk || (s.k_has0 && r == 0)) ? 0 : s.k_base + (long long)(r - s.k_has0) * s.k_stride;
    cp_async16(sQ + swzp<P>(r, c), qb + qt * rs + c * 8, okq);
    cp_async16(sD + swzp<P>(r, c), dob + qt * ro + c * 8, okq);
    cp_async16(sK + swzp<P>(r, c), qb + kt * rs + ro + c * 8, okk);
    cp_async16(sV + swzp<P>(r, c), qb + kt * rs + 2 * ro + c * 8, okk);
  }
  cp_async_commit();
  for (int r = threadIdx.x; r < R; r += NW * 32) {
    const bool ok = r < nq_all;
    const long long tok = ok ? q_token(r) : 0;
    Ls[r] = ok ? lse_b[tok] * LOG2E : 0.f;
    Ds[r] = ok ? delta_b[tok] : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  const float sl2 = a.scale * LOG2E;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;     // this lane's ldmatrix row for A fragments
  const bool own_q = warp * 16 < s.nq;       // this warp owns query rows
  const bool own_k = warp * 16 < nk;         // ... and key rows

  // ---------------- pass A: rows = this warp's queries, columns = keys in chunks of 32: dQ = scale * dS K
  if (own_q) {
    float dq[G::ND][4];
#pragma unroll
    for (int j = 0; j < G::ND; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
    const float l0 = Ls[row0], l1 = Ls[row1], d0 = Ds[row0], d1 = Ds[row1];
    const int k_lim = a.causal ? min(nk, warp * 16 + 16) : nk;
#pragma unroll 1
    for (int kc = 0; kc < k_lim; kc += 32) {
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t qa[4], da[4];       // A fragments are re-read per chunk (2 ldmatrix) instead of living in registers
        ldsm_x4(sQ + swzp<P>(arow, 2 * kk + (lane >> 4)), qa[0], qa[1], qa[2], qa[3]);
        ldsm_x4(sD + swzp<P>(arow, 2 * kk + (lane >> 4)), da[0], da[1], da[2], da[3]);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (kc + 8 * j < k_lim) {
            uint32_t b0, b1, c0, c1;
            g_load_b_rows<HD>(sK, (kc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(sc[j], qa, b0, b1);
            mma16816(sc[j + 1], qa, c0, c1);
            g_load_b_rows<HD>(sV, (kc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(dp[j], da, b0, b1);
            mma16816(dp[j + 1], da, c0, c1);
          }
        }
      }
      uint32_t dsa[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kc + 8 * j + 2 * t4 + e;
          bool ok0 = key < nk && row0 < s.nq, ok1 = key < nk && row1 < s.nq;
          if (a.causal) { ok0 = ok0 && key <= row0; ok1 = ok1 && key <= row1; }
          const float p0 = ok0 ? exp2f(sc[j][e] * sl2 - l0) : 0.f;
          const float p1 = ok1 ? exp2f(sc[j][2 + e] * sl2 - l1) : 0.f;
          ds[e] = p0 * (dp[j][e] - d0);
          ds[2 + e] = p1 * (dp[j][2 + e] - d1);
        }
        dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
        dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (kc + 16 * kk < k_lim) {
#pragma unroll
          for (int jd = 0; jd < G::ND; jd += 2) {
            uint32_t b0, b1, c0, c1;
            g_load_b_cols<HD>(sK, (kc >> 4) + kk, jd, lane, b0, b1, c0, c1);
            mma16816(dq[jd], dsa[kk], b0, b1);
            mma16816(dq[jd + 1], dsa[kk], c0, c1);
          }
        }
      }
    }
    // dq of this warp's rows -> its private rows of the staging tile (frees the registers before pass B)
#pragma unroll
    for (int j = 0; j < G::ND; ++j) {
      *reinterpret_cast<uint32_t*>(gO + swzp<P>(row0, j) + 4 * t4) = pack_bf16x2(dq[j][0] * a.scale, dq[j][1] * a.scale);
      *reinterpret_cast<uint32_t*>(gO + swzp<P>(row1, j) + 4 * t4) = pack_bf16x2(dq[j][2] * a.scale, dq[j][3] * a.scale);
    }
  }
  // ---------------- pass B: rows = this warp's keys, columns = queries (+ CLS query) in chunks of 32: dK = scale * dS^T Q ; dV = P^T dO
  float dk[G::ND][4], dv[G::ND][4];
#pragma unroll
  for (int j = 0; j < G::ND; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  if (own_k) {
    const int q_begin = a.causal ? (warp * 16) & ~31 : 0;      // queries before this warp's first key never see it
#pragma unroll 1
    for (int qc = q_begin; qc < nq_all; qc += 32) {
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t ka[4], va[4];
        ldsm_x4(sK + swzp<P>(arow, 2 * kk + (lane >> 4)), ka[0], ka[1], ka[2], ka[3]);
        ldsm_x4(sV + swzp<P>(arow, 2 * kk + (lane >> 4)), va[0], va[1], va[2], va[3]);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (qc + 8 * j < nq_all) {
            uint32_t b0, b1, c0, c1;
            g_load_b_rows<HD>(sQ, (qc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(sc[j], ka, b0, b1);
            mma16816(sc[j + 1], ka, c0, c1);
            g_load_b_rows<HD>(sD, (qc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(dp[j], va, b0, b1);
            mma16816(dp[j + 1], va, c0, c1);
          }
        }
      }
      uint32_t pa[2][4], dsa[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q = qc + 8 * j + 2 * t4 + e;
          const bool okq = q < nq_all;
          bool ok0 = okq && row0 < nk, ok1 = okq && row1 < nk;
          if (a.causal) { ok0 = ok0 && q >= row0; ok1 = ok1 && q >= row1; }
          const float lq = okq ? Ls[q] : 0.f, dq_ = okq ? Ds[q] : 0.f;
          p[e] = ok0 ? exp2f(sc[j][e] * sl2 - lq) : 0.f;
          p[2 + e] = ok1 ? exp2f(sc[j][2 + e] * sl2 - lq) : 0.f;
          ds[e] = p[e] * (dp[j][e] - dq_);
          ds[2 + e] = p[2 + e] * (dp[j][2 + e] - dq_);
        }
        pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p[0], p[1]); pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
        dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]); dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (qc + 16 * kk < nq_all) {
#pragma unroll
          for (int jd = 0; jd < G::ND; jd += 2) {
            uint32_t b0, b1, c0, c1;
            g_load_b_cols<HD>(sQ, (qc >> 4) + kk, jd, lane, b0, b1, c0, c1);
            mma16816(dk[jd], dsa[kk], b0, b1);
            mma16816(dk[jd + 1], dsa[kk], c0, c1);
            g_load_b_cols<HD>(sD, (qc >> 4) + kk, jd, lane, b0, b1, c0, c1);
            mma16816(dv[jd], pa[kk], b0, b1);
            mma16816(dv[jd + 1], pa[kk], c0, c1);
          }
        }
      }
    }
  }
  // ---------------- outputs: every warp is done reading the tiles -> stage dk / dv in place of this warp's K / V rows
  __syncthreads();
#pragma unroll
  for (int j = 0; j < G::ND; ++j) {
    *reinterpret_cast<uint32_t*>(gsm + R * P + swzp<P>(row0, j) + 4 * t4) = pack_bf16x2(dk[j][0] * a.scale, dk[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(gsm + R * P + swzp<P>(row1, j) + 4 * t4) = pack_bf16x2(dk[j][2] * a.scale, dk[j][3] * a.scale);
    *reinterpret_cast<uint32_t*>(gsm + 2 * R * P + swzp<P>(row0, j) + 4 * t4) = pack_bf16x2(dv[j][0], dv[j][1]);
    *reinterpret_cast<uint32_t*>(gsm + 2 * R * P + swzp<P>(row1, j) + 4 * t4) = pack_bf16x2(dv[j][2], dv[j][3]);
  }
  __syncwarp();
  for (int i = lane; i < 16 * CH; i += 32) {
    const int rr = i / CH, c = i - rr * CH;
    const int r = warp * 16 + rr;
    if (r < s.nq)
      *reinterpret_cast<uint4*>(dq_b + (s.q_base + (long long)r * s.q_stride) * rs + c * 8) = *reinterpret_cast<const uint4*>(gO + swzp<P>(r, c));
    if (r < nk && !(s.k_has0 && r == 0)) {            // key 0 = CLS in space mode: written by the cls_only launch
      const long long kt = s.k_base + (long long)(r - s.k_has0) * s.k_stride;
      *reinterpret_cast<uint4*>(dq_b + kt * rs + ro + c * 8) = *reinterpret_cast<const uint4*>(gsm + R * P + swzp<P>(r, c));
      *reinterpret_cast<uint4*>(dq_b + kt * rs + 2 * ro + c * 8) = *reinterpret_cast<const uint4*>(gsm + 2 * R * P + swzp<P>(r, c));
    }
  }
}


// ================================================================================================ time mode: warp per slot
// One warp owns (batch b, head h, slot j): queries = tokens 1 + f*n + j (f < T), keys = [CLS ; the same T tokens].  The slot's rows fill
// MT 16-row MMA tiles (MT = 1: T <= 15, the shipped 12-frame clips -- same algorithm as attn_time_fwd/bwd_kernel in attention.cu;
// MT = 2: T <= 31, the 16-frame clips of BASELINE.json configs[3]).  Per-warp shared memory, rows of pitch Geo<HD>::PITCH, R = 16*MT:
//   forward : Q [R] | K [R] | V [R]            backward: Q [R] | K [R] | V [R] | dO [R]
// In backward, row T of the Q / dO tiles holds the CLS query (one more column of the transposed pass that produces dK / dV); its own dq
// and the CLS key's dk / dv come from the cls_only streamed launch.
constexpr int HD_TW = 4;            // warps (slots) per CTA

template <int HD, int MT, bool BWD>
__device__ __forceinline__ void hd_time_load(uint32_t base, const bf16* qb, const bf16* dob, long long rs, long long ro, int T, int n, int slot,
                                             int lane) {
  constexpr int P = Geo<HD>::PITCH, CH = Geo<HD>::CH, R = 16 * MT;
  const bf16* cls = qb;                                        // token 0
  for (int i = lane; i < R * CH; i += 32) {
    const int rr = i / CH, c = i - rr * CH;
    const uint32_t so = swzp<P>(rr, c);                        // block bases are multiples of 8 rows: same swizzle in every block
    const bool okq = rr < T, okc = BWD && rr == T, okk = rr <= T;
    const bf16* pq = qb + (1LL + slot + (long long)rr * n) * rs + c * 8;             // query-side row rr: token 1 + rr*n + slot
    const bf16* pk = qb + (1LL + slot + (long long)(rr - 1) * n) * rs + ro + c * 8;  // key-side row rr >= 1: token 1 + (rr-1)*n + slot
    // (rows that are not loaded are zero-filled; they are still given a valid address)
    cp_async16(base + so, okq ? pq : cls + c * 8, okq || okc);                                               // Q  (row T: the CLS query, backward)
    cp_async16(base + R * P + so, (rr == 0 || !okk) ? cls + ro + c * 8 : pk, okk);                           // K  (row 0: the CLS key)
    cp_async16(base + 2 * R * P + so, (rr == 0 || !okk) ? cls + 2 * ro + c * 8 : pk + ro, okk);              // V
    if (BWD) cp_async16(base + 3 * R * P + so, okq ? dob + (1LL + slot + (long long)rr * n) * ro + c * 8 : dob + c * 8, okq || okc);   // dO
  }
}

template <int HD, int MT>
__global__ void __launch_bounds__(HD_TW * 32) attn_hd_time_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                      float* __restrict__ lse, AttnShape a) {
  using G = Geo<HD>;
  constexpr int P = G::PITCH, CH = G::CH, R = 16 * MT, ROWS = 3 * R;
  TVTS_DYN_SMEM(uint8_t, smem, 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = blockIdx.x * HD_TW + warp;
  if (slot >= a.n) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const uint32_t base = smem_u32(smem) + warp * ROWS * P;
  hd_time_load<HD, MT, false>(base, qb, nullptr, rs, ro, a.T, a.n, slot, lane);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
  const uint32_t sQ = base, sK = base + R * P, sV = base + 2 * R * P;
  const int nk = a.T + 1;
  const float sl2 = a.scale * LOG2E;
  uint8_t* stg = smem + warp * ROWS * P;   // the Q rows are reused for output staging (each m-tile's rows after its fragments are loaded)
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
  float* lb = lse + ((long long)b * a.H + h) * a.N;
#pragma unroll
  for (int mi = 0; mi < MT; ++mi) {
    if (16 * mi >= a.T) break;                                   // warp-uniform
    uint32_t qa[G::KS][4];
    g_load_a<HD>(sQ, 16 * mi, lane, qa);
    float sc[2 * MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * MT; j += 2) {
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t b0, b1, c0, c1;
        g_load_b_rows<HD>(sK, j, kk, lane, b0, b1, c0, c1);
        mma16816(sc[j], qa[kk], b0, b1);
        mma16816(sc[j + 1], qa[kk], c0, c1);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (8 * j + 2 * t4 + e >= nk) { sc[j][e] = -INFINITY; sc[j][2 + e] = -INFINITY; }
        mx0 = fmaxf(mx0, sc[j][e]); mx1 = fmaxf(mx1, sc[j][2 + e]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) {
      const float p00 = exp2f((sc[j][0] - mx0) * sl2), p01 = exp2f((sc[j][1] - mx0) * sl2);
      const float p10 = exp2f((sc[j][2] - mx1) * sl2), p11 = exp2f((sc[j][3] - mx1) * sl2);
      l0 += p00 + p01; l1 += p10 + p11;
      pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p00, p01);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p10, p11);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[G::ND][4];
#pragma unroll
    for (int j = 0; j < G::ND; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
#pragma unroll
      for (int jd = 0; jd < G::ND; jd += 2) {
        uint32_t b0, b1, c0, c1;
        g_load_b_cols<HD>(sV, kk, jd, lane, b0, b1, c0, c1);
        mma16816(o[jd], pa[kk], b0, b1);
        mma16816(o[jd + 1], pa[kk], c0, c1);
      }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();
    const int r0 = 16 * mi + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < G::ND; ++j) {
      *reinterpret_cast<uint32_t*>(stg + swzp<P>(r0, j) + 4 * t4) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
      *reinterpret_cast<uint32_t*>(stg + swzp<P>(r1, j) + 4 * t4) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
      const int rr = i / CH, c = i - rr * CH;
      const int r = 16 * mi + rr;
      if (r < a.T) *reinterpret_cast<uint4*>(ob + (long long)(1 + r * a.n + slot) * ro + c * 8) = *reinterpret_cast<const uint4*>(stg + swzp<P>(r, c));
    }
    if (t4 == 0) {
      if (r0 < a.T) lb[1 + r0 * a.n + slot] = mx0 * a.scale + __logf(l0);
      if (r1 < a.T) lb[1 + r1 * a.n + slot] = mx1 * a.scale + __logf(l1);
    }
  }
}

// backward of one slot: dq (T rows) and dk / dv of the slot's T patch keys (including the CLS query's contribution)
template <int HD, int MT>
__global__ void __launch_bounds__(HD_TW * 32) attn_hd_time_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                      const float* __restrict__ lse, const float* __restrict__ delta,
                                                                      bf16* __restrict__ dqkv, AttnShape a) {
  using G = Geo<HD>;
  constexpr int P = G::PITCH, CH = G::CH, R = 16 * MT, ROWS = 4 * R;
  TVTS_DYN_SMEM(uint8_t, smem, 128);                   // per warp: Q | K | V | dO (R rows each); then per warp lse*log2e [R], delta [R]
  float* stat_all = reinterpret_cast<float*>(smem + HD_TW * ROWS * P);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = blockIdx.x * HD_TW + warp;
  if (slot >= a.n) return;
  float* st_l = stat_all + warp * 2 * R;               // lse * log2e of query rows 0..T-1 (slot) and T (CLS)
  float* st_d = st_l + R;                              // delta of the same rows
  const int h = blockIdx.y, b = blockIdx.z;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  const uint32_t base = smem_u32(smem) + warp * ROWS * P;
  hd_time_load<HD, MT, true>(base, qb, dob, rs, ro, a.T, a.n, slot, lane);
  cp_async_commit();
  if (lane < R) {
    const bool ok = lane <= a.T;
    const int tok = lane < a.T ? 1 + lane * a.n + slot : 0;
    st_l[lane] = ok ? lse_b[tok] * LOG2E : 0.f;
    st_d[lane] = ok ? delta_b[tok] : 0.f;
  }
  cp_async_wait<0>();
  __syncwarp();
  const uint32_t sQ = base, sK = base + R * P, sV = base + 2 * R * P, sD = base + 3 * R * P;
  const int nk = a.T + 1, nq_all = a.T + 1;            // keys: CLS + T ; query columns of the transposed pass: T + the CLS query
  const float sl2 = a.scale * LOG2E;
  uint8_t* wsm = smem + warp * ROWS * P;
  bf16* dq_b = dqkv + (long long)b * a.N * rs + (long long)h * HD;

  // Both passes only READ the Q / K / V / dO tiles; their results are staged into the tiles AFTER the last read (dq over Q, dk over K,
  // dv over V), which needs all of dq / dk / dv in registers at once only for MT = 1.  For MT = 2 the m-tiles are processed one at a
  // time and each tile's dq is parked in ITS OWN rows of the dO tile... which pass B still reads -- so dq goes to global memory directly.
  // ---------------- pass A: rows = queries (m-tile mi), cols = keys: dQ = scale * dS K
#pragma unroll
  for (int mi = 0; mi < MT; ++mi) {
    if (16 * mi >= a.T) break;                         // warp-uniform: no slot query in this m-tile (row T = CLS is not this kernel's)
    uint32_t qa[G::KS][4], da[G::KS][4];
    g_load_a<HD>(sQ, 16 * mi, lane, qa);
    g_load_a<HD>(sD, 16 * mi, lane, da);
    float sc[2 * MT][4], dp[2 * MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 2 * MT; j += 2) {
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t b0, b1, c0, c1;
        g_load_b_rows<HD>(sK, j, kk, lane, b0, b1, c0, c1);
        mma16816(sc[j], qa[kk], b0, b1);
        mma16816(sc[j + 1], qa[kk], c0, c1);
        g_load_b_rows<HD>(sV, j, kk, lane, b0, b1, c0, c1);
        mma16816(dp[j], da[kk], b0, b1);
        mma16816(dp[j + 1], da[kk], c0, c1);
      }
    }
    const int r0 = 16 * mi + g, r1 = r0 + 8;
    const float l0 = st_l[r0], l1 = st_l[r1], d0 = st_d[r0], d1 = st_d[r1];
    uint32_t dsa[MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = 8 * j + 2 * t4 + e < nk;
        const float p0 = (ok && r0 < a.T) ? exp2f(sc[j][e] * sl2 - l0) : 0.f;          // row T (CLS) is not this kernel's
        const float p1 = (ok && r1 < a.T) ? exp2f(sc[j][2 + e] * sl2 - l1) : 0.f;
        ds[e] = p0 * (dp[j][e] - d0);
        ds[2 + e] = p1 * (dp[j][2 + e] - d1);
      }
      dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
      dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    float dq[G::ND][4];
#pragma unroll
    for (int j = 0; j < G::ND; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
#pragma unroll
      for (int jd = 0; jd < G::ND; jd += 2) {
        uint32_t b0, b1, c0, c1;
        g_load_b_cols<HD>(sK, kk, jd, lane, b0, b1, c0, c1);
        mma16816(dq[jd], dsa[kk], b0, b1);
        mma16816(dq[jd + 1], dsa[kk], c0, c1);
      }
    }
    // dq rows of this m-tile straight to global memory (4-byte stores; 2 rows x HD/8 chunks per lane)
#pragma unroll
    for (int j = 0; j < G::ND; ++j) {
      if (r0 < a.T)
        *reinterpret_cast<uint32_t*>(dq_b + (1LL + (long long)r0 * a.n + slot) * rs + 8 * j + 2 * t4) = pack_bf16x2(dq[j][0] * a.scale, dq[j][1] * a.scale);
      if (r1 < a.T)
        *reinterpret_cast<uint32_t*>(dq_b + (1LL + (long long)r1 * a.n + slot) * rs + 8 * j + 2 * t4) = pack_bf16x2(dq[j][2] * a.scale, dq[j][3] * a.scale);
    }
  }
  // ---------------- pass B: rows = keys (m-tile mi), cols = queries (slot queries + CLS query): dK = scale * dS^T Q ; dV = P^T dO
#pragma unroll
  for (int mi = 0; mi < MT; ++mi) {
    if (16 * mi >= nk) break;
    float dk[G::ND][4], dv[G::ND][4];
#pragma unroll
    for (int j = 0; j < G::ND; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
    uint32_t ka[G::KS][4], va[G::KS][4];
    g_load_a<HD>(sK, 16 * mi, lane, ka);
    g_load_a<HD>(sV, 16 * mi, lane, va);
    float sc[2 * MT][4], dp[2 * MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 2 * MT; j += 2) {
#pragma unroll
      for (int kk = 0; kk < G::KS; ++kk) {
        uint32_t b0, b1, c0, c1;
        g_load_b_rows<HD>(sQ, j, kk, lane, b0, b1, c0, c1);
        mma16816(sc[j], ka[kk], b0, b1);
        mma16816(sc[j + 1], ka[kk], c0, c1);
        g_load_b_rows<HD>(sD, j, kk, lane, b0, b1, c0, c1);
        mma16816(dp[j], va[kk], b0, b1);
        mma16816(dp[j + 1], va[kk], c0, c1);
      }
    }
    uint32_t pa[MT][4], dsa[MT][4];
    const int key0 = 16 * mi + g, key1 = key0 + 8;
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) {
      float p[4], ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = 8 * j + 2 * t4 + e;
        const bool okq = q < nq_all;         // q == T: the CLS query
        const float lq = st_l[q], dq_ = st_d[q];
        p[e] = (okq && key0 < nk) ? exp2f(sc[j][e] * sl2 - lq) : 0.f;
        p[2 + e] = (okq && key1 < nk) ? exp2f(sc[j][2 + e] * sl2 - lq) : 0.f;
        ds[e] = p[e] * (dp[j][e] - dq_);
        ds[2 + e] = p[2 + e] * (dp[j][2 + e] - dq_);
      }
      pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p[0], p[1]); pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
      dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]); dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
#pragma unroll
      for (int jd = 0; jd < G::ND; jd += 2) {
        uint32_t b0, b1, c0, c1;
        g_load_b_cols<HD>(sQ, kk, jd, lane, b0, b1, c0, c1);
        mma16816(dk[jd], dsa[kk], b0, b1);
        mma16816(dk[jd + 1], dsa[kk], c0, c1);
        g_load_b_cols<HD>(sD, kk, jd, lane, b0, b1, c0, c1);
        mma16816(dv[jd], pa[kk], b0, b1);
        mma16816(dv[jd + 1], pa[kk], c0, c1);
      }
    }
    // dk / dv rows of this key m-tile straight to global memory (key row 0 = CLS: written by the cls_only launch)
#pragma unroll
    for (int j = 0; j < G::ND; ++j) {
      if (key0 >= 1 && key0 < nk) {
        bf16* dst = dq_b + (1LL + (long long)(key0 - 1) * a.n + slot) * rs + 8 * j + 2 * t4;
        *reinterpret_cast<uint32_t*>(dst + ro) = pack_bf16x2(dk[j][0] * a.scale, dk[j][1] * a.scale);
        *reinterpret_cast<uint32_t*>(dst + 2 * ro) = pack_bf16x2(dv[j][0], dv[j][1]);
      }
      if (key1 < nk) {
        bf16* dst = dq_b + (1LL + (long long)(key1 - 1) * a.n + slot) * rs + 8 * j + 2 * t4;
        *reinterpret_cast<uint32_t*>(dst + ro) = pack_bf16x2(dk[j][2] * a.scale, dk[j][3] * a.scale);
        *reinterpret_cast<uint32_t*>(dst + 2 * ro) = pack_bf16x2(dv[j][2], dv[j][3]);
      }
    }
  }
  (void)wsm; (void)CH;
}

// tiles per slot for the warp-per-slot time kernels (0: T too large -> streamed kernels)
__host__ __device__ inline int hd_time_tiles(const AttnShape& a) { return a.mode != 2 ? 0 : (a.T <= 15 ? 1 : (a.T <= 31 ? 2 : 0)); }

// number of warps a group-resident launch needs (0: not applicable -> streamed kernels); same rule as group_warps() in attention.cu
__host__ __device__ inline int hd_group_warps(const AttnShape& a) {
  if (a.mode == 2) return 0;
  const int m = a.mode == 0 ? a.N : a.n + 1;           // space mode: the CLS key / (in backward) the CLS query is one extra row
  const int nw = (m + 15) / 16;
  return (nw >= 2 && nw <= 7) ? nw : 0;
}

// ------------------------------------------------------------------------------------------------ host side
#ifndef TVTS_HOST_SHIM
int hd_check_shape(const AttnShape& a, int64_t d) {
  TVTS_REQUIRE(d == 64 || d == 80, "attention: head dim %lld unsupported (64 and 80 are built)", (long long)d);
  TVTS_REQUIRE(a.B > 0 && a.N > 0 && a.H > 0, "attention: empty shape");
  TVTS_REQUIRE(a.mode >= 0 && a.mode <= 2, "attention: bad mode %d", a.mode);
  if (a.mode != 0) {
    TVTS_REQUIRE(a.T > 0 && a.n > 0 && a.N == 1 + a.T * a.n, "attention: N=%d != 1 + T*n (T=%d n=%d)", a.N, a.T, a.n);
    TVTS_REQUIRE(!a.causal, "attention: causal only valid in full mode");
  }
  TVTS_REQUIRE(a.H <= 65535 && a.B <= 65535, "attention: grid limits");
  return TVTS_OK;
}

int g_hd_group = 1;      // 1: group-resident kernels where a group fits one CTA; 0: streamed kernels only (tvts_attn_hd_set_group)

// The CLS row / column launch of the divided modes is latency-bound and writes outputs disjoint from the group / time kernels', so it
// runs on a side stream forked from / joined back into the caller's stream with events (legal inside a CUDA-graph capture: a parallel
// branch) -- the same arrangement as in attention.cu.
struct HdSideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
HdSideStream* hd_side_stream() {
  static HdSideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HdSideStream& s = per_dev[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}
int g_hd_side = 1;       // tvts_attn_hd_set_side_stream

template <int HD>
int hd_streamed_fwd(const void* qkv, void* out, float* lse, const AttnShape& a, cudaStream_t st, const int* klen) {
  constexpr int smem_bytes = 5 * Geo<HD>::TILE;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 grid(num_blocks_x(a), (unsigned)a.H, (unsigned)a.B);
  attn_hd_fwd_kernel<HD><<<grid, kThreads, smem_bytes, st>>>((const bf16*)qkv, (bf16*)out, lse, a, klen);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD, int NW>
int hd_group_fwd(const void* qkv, void* out, float* lse, const AttnShape& a, cudaStream_t st) {
  constexpr int smem_bytes = 3 * NW * 16 * Geo<HD>::PITCH;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_group_fwd_kernel<HD, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 gg((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B);
  attn_hd_group_fwd_kernel<HD, NW><<<gg, NW * 32, smem_bytes, st>>>((const bf16*)qkv, (bf16*)out, lse, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD, int MT>
int hd_time_fwd(const void* qkv, void* out, float* lse, const AttnShape& a, cudaStream_t st) {
  constexpr int smem_bytes = HD_TW * 3 * 16 * MT * Geo<HD>::PITCH;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_time_fwd_kernel<HD, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 tg((unsigned)((a.n + HD_TW - 1) / HD_TW), (unsigned)a.H, (unsigned)a.B);
  attn_hd_time_fwd_kernel<HD, MT><<<tg, HD_TW * 32, smem_bytes, st>>>((const bf16*)qkv, (bf16*)out, lse, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD, int MT>
int hd_time_bwd(const void* qkv, const void* dout, const float* lse, const float* delta_ws, void* dqkv, const AttnShape& a, cudaStream_t st) {
  constexpr int smem_bytes = HD_TW * 4 * 16 * MT * Geo<HD>::PITCH + HD_TW * 2 * 16 * MT * 4;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_time_bwd_kernel<HD, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 tg((unsigned)((a.n + HD_TW - 1) / HD_TW), (unsigned)a.H, (unsigned)a.B);
  attn_hd_time_bwd_kernel<HD, MT><<<tg, HD_TW * 32, smem_bytes, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD>
int hd_launch_fwd(const void* qkv, void* out, float* lse, const AttnShape& a, cudaStream_t st, const int* klen = nullptr) {
  const int gw = (klen == nullptr && g_hd_group) ? hd_group_warps(a) : 0;
  const int time_k = (klen == nullptr && g_hd_group) ? hd_time_tiles(a) : 0;
  if (!gw && !time_k) return hd_streamed_fwd<HD>(qkv, out, lse, a, st, klen);
  HdSideStream* sd = (a.mode != 0 && g_hd_side) ? hd_side_stream() : nullptr;
  if (a.mode != 0) {                     // the CLS query over all tokens
    cudaStream_t cls_st = st;
    if (sd) {
      TVTS_CHECK_CUDA(cudaEventRecord(sd->fork, st));
      TVTS_CHECK_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
      cls_st = sd->stream;
    }
    AttnShape c = a;
    c.cls_only = 1;
    int rc = hd_streamed_fwd<HD>(qkv, out, lse, c, cls_st, nullptr);
    if (rc) return rc;
  }
  int rc;
  if (time_k) rc = time_k == 1 ? hd_time_fwd<HD, 1>(qkv, out, lse, a, st) : hd_time_fwd<HD, 2>(qkv, out, lse, a, st);
  else
    switch (gw) {
      case 2: rc = hd_group_fwd<HD, 2>(qkv, out, lse, a, st); break;
      case 3: rc = hd_group_fwd<HD, 3>(qkv, out, lse, a, st); break;
      case 4: rc = hd_group_fwd<HD, 4>(qkv, out, lse, a, st); break;
      case 5: rc = hd_group_fwd<HD, 5>(qkv, out, lse, a, st); break;
      case 6: rc = hd_group_fwd<HD, 6>(qkv, out, lse, a, st); break;
      default: rc = hd_group_fwd<HD, 7>(qkv, out, lse, a, st); break;
    }
  if (sd) {                              // join (also on the error path, so that the side branch never dangles in a capture)
    TVTS_CHECK_CUDA(cudaEventRecord(sd->join, sd->stream));
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
  }
  return rc;
}

template <int HD>
int hd_streamed_bwd(const void* qkv, const void* dout, const float* lse, const float* delta_ws, void* dqkv, const AttnShape& a,
                    cudaStream_t st, const int* klen) {
  constexpr int smem_bytes = 4 * Geo<HD>::TILE + 2 * 2 * BN * 4;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_bwd_kernel<HD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_bwd_kernel<HD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 g0(num_blocks_x(a, false), (unsigned)a.H, (unsigned)a.B);
  attn_hd_bwd_kernel<HD, 0><<<g0, kThreads, smem_bytes, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a, klen);
  TVTS_LAUNCH_CHECK();
  dim3 g1(num_blocks_x(a, true), (unsigned)a.H, (unsigned)a.B);
  attn_hd_bwd_kernel<HD, 1><<<g1, kThreads, smem_bytes, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a, klen);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD, int NW>
int hd_group_bwd(const void* qkv, const void* dout, const float* lse, const float* delta_ws, void* dqkv, const AttnShape& a, cudaStream_t st) {
  constexpr int smem_bytes = 5 * NW * 16 * Geo<HD>::PITCH + 2 * NW * 16 * 4;
  static bool set = false;
  if (!set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_hd_group_bwd_kernel<HD, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    set = true;
  }
  dim3 gg((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)a.H, (unsigned)a.B);
  attn_hd_group_bwd_kernel<HD, NW><<<gg, NW * 32, smem_bytes, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

template <int HD>
int hd_launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, const AttnShape& a,
                  cudaStream_t st, const int* klen = nullptr) {
  const long long rows = (long long)a.B * a.N * a.H;
  attn_hd_delta_kernel<HD><<<(unsigned)((rows + 31) / 32), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta_ws, a.B, a.N, a.H);
  TVTS_LAUNCH_CHECK();
  const int gw = (klen == nullptr && g_hd_group) ? hd_group_warps(a) : 0;
  const int time_k = (klen == nullptr && g_hd_group) ? hd_time_tiles(a) : 0;
  if (!gw && !time_k) return hd_streamed_bwd<HD>(qkv, dout, lse, delta_ws, dqkv, a, st, klen);
  HdSideStream* sd = (a.mode != 0 && g_hd_side) ? hd_side_stream() : nullptr;
  if (a.mode != 0) {                     // dq of the CLS query, dk / dv of the CLS key
    cudaStream_t cls_st = st;
    if (sd) {
      TVTS_CHECK_CUDA(cudaEventRecord(sd->fork, st));      // after delta
      TVTS_CHECK_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
      cls_st = sd->stream;
    }
    AttnShape c = a;
    c.cls_only = 1;
    int rc = hd_streamed_bwd<HD>(qkv, dout, lse, delta_ws, dqkv, c, cls_st, nullptr);
    if (rc) return rc;
  }
  int rc;
  if (time_k) rc = time_k == 1 ? hd_time_bwd<HD, 1>(qkv, dout, lse, delta_ws, dqkv, a, st) : hd_time_bwd<HD, 2>(qkv, dout, lse, delta_ws, dqkv, a, st);
  else
    switch (gw) {
      case 2: rc = hd_group_bwd<HD, 2>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
      case 3: rc = hd_group_bwd<HD, 3>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
      case 4: rc = hd_group_bwd<HD, 4>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
      case 5: rc = hd_group_bwd<HD, 5>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
      case 6: rc = hd_group_bwd<HD, 6>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
      default: rc = hd_group_bwd<HD, 7>(qkv, dout, lse, delta_ws, dqkv, a, st); break;
    }
  if (sd) {
    TVTS_CHECK_CUDA(cudaEventRecord(sd->join, sd->stream));
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
  }
  return rc;
}

#endif  // !TVTS_HOST_SHIM

}  // namespace

#ifndef TVTS_HOST_SHIM
extern "C" int tvts_attn_hd_set_group(int on) {
  g_hd_group = on;
  return TVTS_OK;
}

extern "C" int tvts_attn_hd_set_side_stream(int on) {
  g_hd_side = on;
  return TVTS_OK;
}

extern "C" int tvts_attn_generic_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode,
                                     int64_t T, int64_t n, int64_t causal, float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = hd_check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && lse, "attn_generic_fwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d == 64 ? hd_launch_fwd<64>(qkv, out, lse, a, st) : hd_launch_fwd<80>(qkv, out, lse, a, st);
}

extern "C" int tvts_attn_generic_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                     int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal,
                                     float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = hd_check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv, "attn_generic_bwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d == 64 ? hd_launch_bwd<64>(qkv, out, dout, lse, delta_ws, dqkv, a, st)
                 : hd_launch_bwd<80>(qkv, out, dout, lse, delta_ws, dqkv, a, st);
}

// Key-padded full attention (the v1 text encoder, DistilBERT: padded key positions of `attention_mask` are excluded from the softmax;
// v1/model/model_dist_TVTS.py:124-126).  klen [B] int32 = number of leading valid tokens per sequence (1 <= klen[b] <= N; the tokenizer
// pads on the right).  Query rows >= klen[b] are still computed (over the valid keys), exactly like the reference.
extern "C" int tvts_attn_padded_fwd(const void* qkv, void* out, float* lse, const int32_t* klen, int64_t B, int64_t N, int64_t H, int64_t d,
                                    float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, 0, 0, 0, 0, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = hd_check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && lse && klen, "attn_padded_fwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d == 64 ? hd_launch_fwd<64>(qkv, out, lse, a, st, klen) : hd_launch_fwd<80>(qkv, out, lse, a, st, klen);
}

extern "C" int tvts_attn_padded_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                    const int32_t* klen, int64_t B, int64_t N, int64_t H, int64_t d, float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, 0, 0, 0, 0, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = hd_check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv && klen, "attn_padded_bwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d == 64 ? hd_launch_bwd<64>(qkv, out, dout, lse, delta_ws, dqkv, a, st, klen)
                 : hd_launch_bwd<80>(qkv, out, dout, lse, delta_ws, dqkv, a, st, klen);
}
#endif  // !TVTS_HOST_SHIM
