// Grouped multi-head attention for the TVTS hot path, head dim 64, forward + backward:
//   mode FULL  : every token attends to every token (sort head, v2/model/sort_transformer.py:9-13,43-57), optionally
//                causal (CLIP text tower, v2/CLIP/clip/model.py:185-188 with the mask of :330-336)
//   mode SPACE : divided space attention of VarAttention (v2/model/video_encoder_ViT_B_16.py:38-76, '(b f) n d'):
//                a patch token attends to [CLS ; the n kept tokens of its own frame]; CLS attends to all N tokens
//   mode TIME  : divided time attention ('(b n) f d'): a patch token attends to [CLS ; the T tokens of its slot]
// The reference materialises q_/k_/v_/cls_k/cls_v copies with rearrange/repeat/cat and an eager softmax; here the grouping
// is pure index arithmetic on the packed qkv buffer [B, N, 3, H, 64] written by the qkv GEMM, so nothing is copied and the
// score matrix never leaves registers.
//
// These kernels are HBM-bound (attention is < 4 % of the path's FLOPs): sequences are 9 .. 789 tokens of head dim 64, so
// the math runs on warp-level tensor-core MMAs (mma.sync m16n8k16 bf16, fp32 accumulate) fed by ldmatrix from XOR-swizzled
// shared-memory tiles that are filled with 16-byte cp.async copies (fully coalesced 128-byte rows).
//   * "streamed" kernels (FULL, SPACE, and the CLS row/column of SPACE/TIME): a CTA owns 64 STATIONARY tokens (16 per warp) of
//     one (batch, head, group) and streams the group's other side through double-buffered 64-row tiles, flash-attention style:
//         forward / dQ : stationary = queries, streamed = keys (+values)
//         dK,dV        : stationary = keys,    streamed = queries (+dO)     (same index sets: the relation is symmetric)
//   * "time" kernels: one WARP per (batch, head, slot): T <= 15 queries x (T+1) keys fit one 16x16 MMA tile; forward and the
//     whole backward (dq, dk, dv of the slot) each run in a single pass (more frames fall back to the streamed kernels).
// fp32 softmax statistics (online max/sum in forward; saved log-sum-exp in backward), bf16 I/O.
#include "attention_common.cuh"
#ifndef TVTS_HOST_SHIM
#include "../../include/tvts_b200.h"
#endif

namespace {

constexpr int HD = 64;          // head dim
constexpr int TILE_BYTES = 64 * 128;   // 64 rows x 64 bf16

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows][128 B] tile; the XOR keeps ldmatrix conflict-free
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + (((c ^ r) & 7) << 4)); }

// A fragments (16 rows starting at row0, all 64 dims = 4 k-steps) of a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, int lane, uint32_t a[4][4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(tile + swz(r, 2 * kk + (lane >> 4)), a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}
// B fragments with the tile's ROWS as the n index (S = A . tile^T): n-tiles j, j+1 (rows 8j..8j+15), k-step kk (dims 16kk..)
__device__ __forceinline__ void load_b_rows(uint32_t tile, int j, int kk, int lane, uint32_t& b0, uint32_t& b1, uint32_t& c0, uint32_t& c1) {
  const int r = 8 * j + (lane & 7) + ((lane >> 4) & 1) * 8;
  ldsm_x4(tile + swz(r, 2 * kk + ((lane >> 3) & 1)), b0, b1, c0, c1);
}
// B fragments with the tile's ROWS as the k index (O = P . tile): k-step kk (rows 16kk..16kk+15), n-tiles jd, jd+1 (dims 8jd..)
__device__ __forceinline__ void load_b_cols(uint32_t tile, int kk, int jd, int lane, uint32_t& b0, uint32_t& b1, uint32_t& c0, uint32_t& c1) {
  const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
  ldsm_x4_t(tile + swz(r, jd + (lane >> 4)), b0, b1, c0, c1);
}

// cooperative 64-row tile load (rows >= cnt are zero-filled); token(r) gives the token index of tile row r
template <typename TokFn>
__device__ __forceinline__ void load_tile(uint32_t tile, const bf16* base, long long row_stride, int cnt, TokFn token) {
  for (int i = threadIdx.x; i < 64 * 8; i += kThreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < cnt;
    const long long tok = ok ? token(r) : 0;
    cp_async16(tile + swz(r, c), base + tok * row_stride + c * 8, ok);
  }
}

// ================================================================================================ streamed forward
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                                            AttnShape a) {
  __shared__ __align__(128) uint8_t smem[5 * TILE_BYTES];   // Q | K0 | V0 | K1 | V1
  const uint32_t sQ = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const Sets s = decode_sets(a, blockIdx.x);
  const long long rs = 3LL * a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + (long long)a.H * HD;
  const bf16* vb = kb + (long long)a.H * HD;
  const int total = s.sm_has0 + s.sm_count;
  int t_end = (total + BN - 1) / BN;
  if (a.causal) t_end = min(t_end, (s.st_base + (s.st_count - 1) * s.st_stride) / BN + 1);

  load_tile(sQ, qb, rs, s.st_count, [&](int r) { return s.st_base + r * s.st_stride; });
  auto issue = [&](int t) {
    const uint32_t sK = sQ + (1 + 2 * (t & 1)) * TILE_BYTES, sV = sK + TILE_BYTES;
    const int k0 = t * BN, cnt = min(BN, total - k0);
    load_tile(sK, kb, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    load_tile(sV, vb, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
  };
  issue(0);
  cp_async_commit();

  const bool warp_active = warp * 16 < s.st_count;   // warp-uniform
  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl2 = a.scale * LOG2E;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int qtok0 = s.st_base + row0 * s.st_stride, qtok1 = s.st_base + row1 * s.st_stride;

  for (int t = 0; t < t_end; ++t) {
    if (t + 1 < t_end) issue(t + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) load_a_frags(sQ, warp * 16, lane, qa);
    if (warp_active) {
      const uint32_t sK = sQ + (1 + 2 * (t & 1)) * TILE_BYTES, sV = sK + TILE_BYTES;
      const int k0 = t * BN, cnt = min(BN, total - k0);
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        if (8 * j < cnt) {     // CTA-uniform: key tiles past the end of a short group are never computed
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t b0, b1, c0, c1;
            load_b_rows(sK, j, kk, lane, b0, b1, c0, c1);
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
      for (int j = 0; j < 8; ++j) { o[j][0] *= corr0; o[j][1] *= corr0; o[j][2] *= corr1; o[j][3] *= corr1; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (16 * kk < cnt) {
#pragma unroll
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t b0, b1, c0, c1;
            load_b_cols(sV, kk, jd, lane, b0, b1, c0, c1);
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
  for (int j = 0; j < 8; ++j) {
    const uint32_t w0 = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0), w1 = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    *reinterpret_cast<uint32_t*>(smem + swz(row0, j) + 4 * t4) = w0;
    *reinterpret_cast<uint32_t*>(smem + swz(row1, j) + 4 * t4) = w1;
  }
  __syncwarp();
  const long long ro = (long long)a.H * HD;
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 16 + i * 4 + (lane >> 3), c = lane & 7;
    if (r < s.st_count) {
      const uint4 v = *reinterpret_cast<const uint4*>(smem + swz(r, c));
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
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                                                         int B, int N, int H) {
  // 8 lanes per (token, head) row of 64 bf16: one 16-byte load of O and of dO per lane, reduce inside the octet
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool ok = w < (long long)B * N * H;
  float s = 0.f;
  if (ok) {
    const uint4 o = reinterpret_cast<const uint4*>(out + w * HD)[sub];
    const uint4 d = reinterpret_cast<const uint4*>(dout + w * HD)[sub];
    const uint32_t ov[4] = {o.x, o.y, o.z, o.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = unpack_bf16x2(ov[i]), c = unpack_bf16x2(dv[i]);
      s = fmaf(a.x, c.x, s);
      s = fmaf(a.y, c.y, s);
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

// ================================================================================================ streamed backward
// ROLE 0: stationary = query i (tiles Q, dO), streamed X = K, Y = V        -> dq_i = scale * sum_j ds_ij k_j
// ROLE 1: stationary = key j   (tiles K, V),  streamed X = Q, Y = dO       -> dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i p_ij do_i
// with p_ij = exp(scale q_i.k_j - lse_i), ds_ij = p_ij (do_i.v_j - delta_i).  In ROLE 1 the register tiles hold the TRANSPOSED
// score matrix (rows = keys), so P^T / dS^T are directly the A operands of the dV / dK products.
template <int ROLE>
__global__ void __launch_bounds__(kThreads) attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                            const float* __restrict__ lse, const float* __restrict__ delta,
                                                            bf16* __restrict__ dqkv, AttnShape a) {
  __shared__ __align__(128) uint8_t smem[4 * TILE_BYTES + 2 * 2 * BN * 4];   // A-side(2 tiles; reused as stages) | X0 Y0 | ... see below
  // layout: [0] stationary tile 0 (Q or K) -> after fragment load reused as X stage 1
  //         [1] stationary tile 1 (dO or V) -> reused as Y stage 1
  //         [2] X stage 0, [3] Y stage 0 ; then lse/delta of the streamed rows (ROLE 1) for 2 stages
  const uint32_t s0 = smem_u32(smem);
  float* LsDs = reinterpret_cast<float*>(smem + 4 * TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const Sets s = decode_sets(a, blockIdx.x, ROLE == 1);
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
  // stage buffers: stage 0 -> tiles 2,3 ; stage 1 -> tiles 0,1 (free once the stationary fragments are in registers)
  auto issue = [&](int t, int stage) {
    const uint32_t sX = s0 + (stage == 0 ? 2 : 0) * TILE_BYTES, sY = sX + TILE_BYTES;
    const int k0 = t * BN, cnt = min(BN, total - k0);
    load_tile(sX, xs_base, rs, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    load_tile(sY, ys_base, ys_stride, cnt, [&](int r) { return streamed_token(s, k0 + r); });
    if (ROLE == 1) {
      for (int i = threadIdx.x; i < BN; i += kThreads) {
        float l_ = 0.f, d_ = 0.f;
        if (i < cnt) { const int tok = streamed_token(s, k0 + i); l_ = lse_b[tok]; d_ = delta_b[tok]; }
        LsDs[stage * 2 * BN + i] = l_ * LOG2E;
        LsDs[stage * 2 * BN + BN + i] = d_;
      }
    }
  };
  // stationary tiles
  if (ROLE == 0) {
    load_tile(s0, qb, rs, s.st_count, st_tok);
    load_tile(s0 + TILE_BYTES, dob, ro, s.st_count, st_tok);
  } else {
    load_tile(s0, kb, rs, s.st_count, st_tok);
    load_tile(s0 + TILE_BYTES, vb, rs, s.st_count, st_tok);
  }
  issue(t_begin, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t fa[4][4], fb[4][4];   // ROLE 0: Q, dO ; ROLE 1: K, V   (A operands, 16 rows of this warp)
  load_a_frags(s0, warp * 16, lane, fa);
  load_a_frags(s0 + TILE_BYTES, warp * 16, lane, fb);
  __syncthreads();               // tiles 0,1 may now be overwritten (stage 1)

  const bool warp_active = warp * 16 < s.st_count;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int stok0 = st_tok(row0), stok1 = st_tok(row1);
  float lse0 = 0.f, lse1 = 0.f, dl0 = 0.f, dl1 = 0.f;
  if (ROLE == 0) {
    if (row0 < s.st_count) { lse0 = lse_b[stok0] * LOG2E; dl0 = delta_b[stok0]; }
    if (row1 < s.st_count) { lse1 = lse_b[stok1] * LOG2E; dl1 = delta_b[stok1]; }
  }
  const float sl2 = a.scale * LOG2E;
  float acc0[8][4], acc1[8][4];   // ROLE 0: acc0 = dQ ; ROLE 1: acc0 = dK, acc1 = dV
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc0[j][0] = acc0[j][1] = acc0[j][2] = acc0[j][3] = 0.f;
    if (ROLE == 1) acc1[j][0] = acc1[j][1] = acc1[j][2] = acc1[j][3] = 0.f;
  }

  for (int t = t_begin; t < t_end; ++t) {
    const int stage = (t - t_begin) & 1;
    if (t + 1 < t_end) issue(t + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (warp_active) {
      const uint32_t sX = s0 + (stage == 0 ? 2 : 0) * TILE_BYTES, sY = sX + TILE_BYTES;
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
            for (int kk = 0; kk < 4; ++kk) {
              uint32_t b0, b1, c0, c1;
              load_b_rows(sX, half * 4 + j, kk, lane, b0, b1, c0, c1);
              mma16816(sc[j], fa[kk], b0, b1);
              mma16816(sc[j + 1], fa[kk], c0, c1);
              load_b_rows(sY, half * 4 + j, kk, lane, b0, b1, c0, c1);
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
            if (ROLE == 1) { pa[j >> 1][(j & 1) * 2] = 0u; pa[j >> 1][(j & 1) * 2 + 1] = 0u; }
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
            const float la = ROLE == 0 ? lse0 : Ls[col], lb = ROLE == 0 ? lse1 : Ls[col];
            const float da = ROLE == 0 ? dl0 : Ds[col], db = ROLE == 0 ? dl1 : Ds[col];
            p[e] = ok0 ? exp2f(sc[j][e] * sl2 - la) : 0.f;
            p[2 + e] = ok1 ? exp2f(sc[j][2 + e] * sl2 - lb) : 0.f;
            ds[e] = p[e] * (dp[j][e] - da);
            ds[2 + e] = p[2 + e] * (dp[j][2 + e] - db);
          }
          dsa[j >> 1][(j & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
          dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
          if (ROLE == 1) {
            pa[j >> 1][(j & 1) * 2] = pack_bf16x2(p[0], p[1]);
            pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
          }
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (32 * half + 16 * kk >= cnt) break;
#pragma unroll
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t b0, b1, c0, c1;
            load_b_cols(sX, half * 2 + kk, jd, lane, b0, b1, c0, c1);     // ROLE 0: K (dQ += dS K) ; ROLE 1: Q (dK += dS^T Q)
            mma16816(acc0[jd], dsa[kk], b0, b1);
            mma16816(acc0[jd + 1], dsa[kk], c0, c1);
            if (ROLE == 1) {
              load_b_cols(sY, half * 2 + kk, jd, lane, b0, b1, c0, c1);   // dO (dV += P^T dO)
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
  uint8_t* stg = smem + 2 * TILE_BYTES;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(stg + swz(row0, j) + 4 * t4) = pack_bf16x2(acc0[j][0] * a.scale, acc0[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(stg + swz(row1, j) + 4 * t4) = pack_bf16x2(acc0[j][2] * a.scale, acc0[j][3] * a.scale);
    if (ROLE == 1) {
      *reinterpret_cast<uint32_t*>(stg + TILE_BYTES + swz(row0, j) + 4 * t4) = pack_bf16x2(acc1[j][0], acc1[j][1]);
      *reinterpret_cast<uint32_t*>(stg + TILE_BYTES + swz(row1, j) + 4 * t4) = pack_bf16x2(acc1[j][2], acc1[j][3]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 16 + i * 4 + (lane >> 3), c = lane & 7;
    if (r < s.st_count) {
      bf16* dst = dq_b + (long long)st_tok(r) * rs + (ROLE == 0 ? 0 : ro) + c * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stg + swz(r, c));
      if (ROLE == 1) *reinterpret_cast<uint4*>(dst + ro) = *reinterpret_cast<const uint4*>(stg + TILE_BYTES + swz(r, c));
    }
  }
}

// ================================================================================================ time mode: warp per slot
// One warp owns (batch b, head h, slot j): queries = tokens 1 + f*n + j (f < T <= 15), keys = [CLS ; the same T tokens]
// (T+1 <= 16: one 16x16 score tile).  Per-warp shared memory, 128 B rows, swizzled:
//   forward : Q [16] | K [16] | V [16]            backward: Q [16] | K [16] | V [16] | dO [16]
// In backward, row T of the Q / dO tiles holds the CLS query: it also attends to the slot's keys, so it is one more
// column of the transposed pass that produces dK / dV (its own dq is produced by the cls_only streamed launch).
constexpr int TW = 4;                         // warps (slots) per CTA
constexpr int T_ROWS_FWD = 48;
constexpr int T_ROWS_BWD = 64;
constexpr int T_MAX = 15;

template <bool BWD>
__device__ __forceinline__ void time_load(uint32_t base, const bf16* qb, const bf16* dob, long long rs, long long ro, int T, int n, int slot,
                                          int lane) {
  // lane -> (16-byte chunk c, row r0 + 4*it) of each 16-row block; token pointers advance by a constant 4*n rows per iteration
  const int c = lane & 7, r0 = lane >> 3;
  const long long step_q = 4LL * n * rs, step_o = 4LL * n * ro;
  const bf16* pq = qb + (1LL + slot + (long long)r0 * n) * rs + c * 8;          // query-side row rr: token 1 + rr*n + slot
  const bf16* pk = qb + (1LL + slot + (long long)(r0 - 1) * n) * rs + ro + c * 8;  // key-side row rr >= 1: token 1 + (rr-1)*n + slot
  const bf16* pd = BWD ? dob + (1LL + slot + (long long)r0 * n) * ro + c * 8 : nullptr;
  const bf16* cls = qb + c * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = r0 + 4 * it;
    const uint32_t so = (uint32_t)(rr * 128 + (((c ^ rr) & 7) << 4));         // block bases are multiples of 8 rows: same swizzle
    const bool okq = rr < T, okc = BWD && rr == T, okk = rr <= T;
    // (rows that are not loaded are zero-filled; they are still given a valid address)
    cp_async16(base + so, okq ? pq : cls, okq || okc);                                        // Q   (row T: the CLS query, backward)
    cp_async16(base + 16 * 128 + so, (rr == 0 || !okk) ? cls + ro : pk, okk);                 // K   (row 0: the CLS key)
    cp_async16(base + 32 * 128 + so, (rr == 0 || !okk) ? cls + 2 * ro : pk + ro, okk);        // V
    if (BWD) cp_async16(base + 48 * 128 + so, okq ? pd : dob + c * 8, okq || okc);            // dO
    pq += step_q; pk += step_q;
    if (BWD) pd += step_o;
  }
}

__global__ void __launch_bounds__(TW * 32) attn_time_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                                                AttnShape a) {
  __shared__ __align__(128) uint8_t smem[TW * T_ROWS_FWD * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = blockIdx.x * TW + warp;
  if (slot >= a.n) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const uint32_t base = smem_u32(smem) + warp * T_ROWS_FWD * 128;
  time_load<false>(base, qb, nullptr, rs, ro, a.T, a.n, slot, lane);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
  const uint32_t sQ = base, sK = base + 16 * 128, sV = base + 32 * 128;
  const int nk = a.T + 1;
  uint32_t qa[4][4];
  load_a_frags(sQ, 0, lane, qa);
  float sc[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t b0, b1, c0, c1;
    load_b_rows(sK, 0, kk, lane, b0, b1, c0, c1);
    mma16816(sc[0], qa[kk], b0, b1);
    mma16816(sc[1], qa[kk], c0, c1);
  }
  const float sl2 = a.scale * LOG2E;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (8 * j + 2 * t4 + e >= nk) { sc[j][e] = -INFINITY; sc[j][2 + e] = -INFINITY; }
      mx0 = fmaxf(mx0, sc[j][e]); mx1 = fmaxf(mx1, sc[j][2 + e]);
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float p00 = exp2f((sc[j][0] - mx0) * sl2), p01 = exp2f((sc[j][1] - mx0) * sl2);
    const float p10 = exp2f((sc[j][2] - mx1) * sl2), p11 = exp2f((sc[j][3] - mx1) * sl2);
    l0 += p00 + p01; l1 += p10 + p11;
    pa[j * 2] = pack_bf16x2(p00, p01);
    pa[j * 2 + 1] = pack_bf16x2(p10, p11);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int jd = 0; jd < 8; jd += 2) {
    uint32_t b0, b1, c0, c1;
    load_b_cols(sV, 0, jd, lane, b0, b1, c0, c1);
    mma16816(o[jd], pa, b0, b1);
    mma16816(o[jd + 1], pa, c0, c1);
  }
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __syncwarp();
  uint8_t* stg = smem + warp * T_ROWS_FWD * 128;   // reuse the Q rows
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(stg + swz(g, j) + 4 * t4) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(stg + swz(g + 8, j) + 4 * t4) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
  __syncwarp();
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 4 + (lane >> 3), c = lane & 7;
    if (r < a.T) *reinterpret_cast<uint4*>(ob + (long long)(1 + r * a.n + slot) * ro + c * 8) = *reinterpret_cast<const uint4*>(stg + swz(r, c));
  }
  if (t4 == 0) {
    float* lb = lse + ((long long)b * a.H + h) * a.N;
    if (g < a.T) lb[1 + g * a.n + slot] = mx0 * a.scale + __logf(l0);
    if (g + 8 < a.T) lb[1 + (g + 8) * a.n + slot] = mx1 * a.scale + __logf(l1);
  }
}

// backward of one slot: dq (T rows) and dk/dv of the slot's T patch keys (including the CLS query's contribution).  The CLS
// key's dk/dv and the CLS query's dq sum over ALL tokens and are produced by the streamed kernels launched with cls_only = 1.
__global__ void __launch_bounds__(TW * 32) attn_time_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                const float* __restrict__ lse, const float* __restrict__ delta,
                                                                bf16* __restrict__ dqkv, AttnShape a) {
  __shared__ __align__(128) uint8_t smem[TW * T_ROWS_BWD * 128];
  __shared__ float stat[TW][2][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = blockIdx.x * TW + warp;
  if (slot >= a.n) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  const uint32_t base = smem_u32(smem) + warp * T_ROWS_BWD * 128;
  time_load<true>(base, qb, dob, rs, ro, a.T, a.n, slot, lane);
  cp_async_commit();
  if (lane < 16) {                       // lse / delta of query rows 0..T-1 (slot) and T (CLS)
    const bool ok = lane <= a.T;
    const int tok = lane < a.T ? 1 + lane * a.n + slot : 0;
    stat[warp][0][lane] = ok ? lse_b[tok] * LOG2E : 0.f;
    stat[warp][1][lane] = ok ? delta_b[tok] : 0.f;
  }
  cp_async_wait<0>();
  __syncwarp();
  const uint32_t sQ = base, sK = base + 16 * 128, sV = base + 32 * 128, sD = base + 48 * 128;
  const int nk = a.T + 1;
  const float sl2 = a.scale * LOG2E;
  uint8_t* wsm = smem + warp * T_ROWS_BWD * 128;
  bf16* dq_b = dqkv + (long long)b * a.N * rs + (long long)h * HD;

  // ---------------- pass A: rows = queries, cols = keys: dQ = scale * dS K
  float dq[8][4];
  {
    uint32_t qa[4][4], da[4][4];
    load_a_frags(sQ, 0, lane, qa);
    load_a_frags(sD, 0, lane, da);
    float sc[2][4], dp[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, c0, c1;
      load_b_rows(sK, 0, kk, lane, b0, b1, c0, c1);
      mma16816(sc[0], qa[kk], b0, b1);
      mma16816(sc[1], qa[kk], c0, c1);
      load_b_rows(sV, 0, kk, lane, b0, b1, c0, c1);
      mma16816(dp[0], da[kk], b0, b1);
      mma16816(dp[1], da[kk], c0, c1);
    }
    const float l0 = stat[warp][0][g], l1 = stat[warp][0][g + 8], d0 = stat[warp][1][g], d1 = stat[warp][1][g + 8];
    uint32_t dsa[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = 8 * j + 2 * t4 + e < nk;
        const float p0 = (ok && g < a.T) ? exp2f(sc[j][e] * sl2 - l0) : 0.f;          // row T (CLS) is not this kernel's
        const float p1 = (ok && g + 8 < a.T) ? exp2f(sc[j][2 + e] * sl2 - l1) : 0.f;
        ds[e] = p0 * (dp[j][e] - d0);
        ds[2 + e] = p1 * (dp[j][2 + e] - d1);
      }
      dsa[j * 2] = pack_bf16x2(ds[0], ds[1]);
      dsa[j * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int jd = 0; jd < 8; jd += 2) {
      uint32_t b0, b1, c0, c1;
      load_b_cols(sK, 0, jd, lane, b0, b1, c0, c1);
      mma16816(dq[jd], dsa, b0, b1);
      mma16816(dq[jd + 1], dsa, c0, c1);
    }
  }
  // ---------------- pass B: rows = keys, cols = queries (slot queries + CLS query): dK = scale * dS^T Q ; dV = P^T dO
  float dk[8][4], dv[8][4];
  {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
    uint32_t ka[4][4], va[4][4];
    load_a_frags(sK, 0, lane, ka);
    load_a_frags(sV, 0, lane, va);
    float sc[2][4], dp[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, c0, c1;
      load_b_rows(sQ, 0, kk, lane, b0, b1, c0, c1);
      mma16816(sc[0], ka[kk], b0, b1);
      mma16816(sc[1], ka[kk], c0, c1);
      load_b_rows(sD, 0, kk, lane, b0, b1, c0, c1);
      mma16816(dp[0], va[kk], b0, b1);
      mma16816(dp[1], va[kk], c0, c1);
    }
    uint32_t pa[4], dsa[4];
    const int key0 = g, key1 = g + 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float p[4], ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = 8 * j + 2 * t4 + e;
        const bool okq = q <= a.T;           // q == T: the CLS query
        const float lq = stat[warp][0][q], dq_ = stat[warp][1][q];
        p[e] = (okq && key0 < nk) ? exp2f(sc[j][e] * sl2 - lq) : 0.f;
        p[2 + e] = (okq && key1 < nk) ? exp2f(sc[j][2 + e] * sl2 - lq) : 0.f;
        ds[e] = p[e] * (dp[j][e] - dq_);
        ds[2 + e] = p[2 + e] * (dp[j][2 + e] - dq_);
      }
      pa[j * 2] = pack_bf16x2(p[0], p[1]); pa[j * 2 + 1] = pack_bf16x2(p[2], p[3]);
      dsa[j * 2] = pack_bf16x2(ds[0], ds[1]); dsa[j * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
#pragma unroll
    for (int jd = 0; jd < 8; jd += 2) {
      uint32_t b0, b1, c0, c1;
      load_b_cols(sQ, 0, jd, lane, b0, b1, c0, c1);
      mma16816(dk[jd], dsa, b0, b1);
      mma16816(dk[jd + 1], dsa, c0, c1);
      load_b_cols(sD, 0, jd, lane, b0, b1, c0, c1);
      mma16816(dv[jd], pa, b0, b1);
      mma16816(dv[jd + 1], pa, c0, c1);
    }
  }
  // ---------------- outputs: stage in this warp's smem (all operand reads are done), 16-byte row stores
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(wsm + swz(g, j) + 4 * t4) = pack_bf16x2(dq[j][0] * a.scale, dq[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(wsm + swz(g + 8, j) + 4 * t4) = pack_bf16x2(dq[j][2] * a.scale, dq[j][3] * a.scale);
    *reinterpret_cast<uint32_t*>(wsm + swz(16 + g, j) + 4 * t4) = pack_bf16x2(dk[j][0] * a.scale, dk[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(wsm + swz(24 + g, j) + 4 * t4) = pack_bf16x2(dk[j][2] * a.scale, dk[j][3] * a.scale);
    *reinterpret_cast<uint32_t*>(wsm + swz(32 + g, j) + 4 * t4) = pack_bf16x2(dv[j][0], dv[j][1]);
    *reinterpret_cast<uint32_t*>(wsm + swz(40 + g, j) + 4 * t4) = pack_bf16x2(dv[j][2], dv[j][3]);
  }
  __syncwarp();
  for (int i = lane; i < 48 * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    const int blk = r >> 4, rr = r & 15;     // 0 = dq, 1 = dk, 2 = dv
    long long tok;
    if (blk == 0) { if (rr >= a.T) continue; tok = 1 + (long long)rr * a.n + slot; }
    else { if (rr == 0 || rr > a.T) continue; tok = 1 + (long long)(rr - 1) * a.n + slot; }   // key 0 = CLS: cls_only launch
    *reinterpret_cast<uint4*>(dq_b + tok * rs + blk * ro + c * 8) = *reinterpret_cast<const uint4*>(wsm + swz(r, c));
  }
}

// ================================================================================================ group-resident kernels
// Short groups (space attention: n + 1 <= 112 keys; the 77-token causal text sequences): ONE CTA holds a whole (batch, head,
// group) in shared memory -- every tile is loaded exactly once, there is no streaming loop, no online-softmax rescaling and no
// per-tile barriers; warp w owns query rows (and, in the backward's transposed pass, key rows) 16w .. 16w+15.
//   forward : Q | K | V                     (NW x 16 rows of 128 B each)
//   backward: Q | K | V | dO  (+ lse, delta) -> dq, dk, dv of the group in one launch.  In space mode the CLS query (which attends
//             to every key) is appended as one more row of the Q / dO tiles so that the transposed pass sees its column; the CLS
//             key's dk/dv and the CLS query's dq sum over ALL tokens and come from the cls_only streamed launch.
struct GroupSets {
  int q_base, q_stride, nq;      // queries owned by the group
  int k_has0, k_base, k_stride, nk_patch;   // keys: [token 0 if k_has0] ; nk_patch tokens
};
__device__ __forceinline__ GroupSets group_sets(const AttnShape& a, int g) {
  GroupSets s;
  if (a.mode == 0) { s.q_base = 0; s.q_stride = 1; s.nq = a.N; s.k_has0 = 0; s.k_base = 0; s.k_stride = 1; s.nk_patch = a.N; }
  else { s.q_base = 1 + g * a.n; s.q_stride = 1; s.nq = a.n; s.k_has0 = 1; s.k_base = 1 + g * a.n; s.k_stride = 1; s.nk_patch = a.n; }
  return s;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 2) attn_group_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                     float* __restrict__ lse, AttnShape a) {
  constexpr int R = NW * 16;
  __shared__ __align__(128) uint8_t smem[3 * R * 128];
  const uint32_t sQ = smem_u32(smem), sK = sQ + R * 128, sV = sK + R * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const GroupSets s = group_sets(a, blockIdx.x);
  const int nk = s.k_has0 + s.nk_patch;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  for (int i = threadIdx.x; i < R * 8; i += NW * 32) {
    const int r = i >> 3, c = i & 7;
    const bool okq = r < s.nq, okk = r < nk;
    const long long qt = okq ? s.q_base + (long long)r * s.q_stride : 0;
    const long long kt = (!okk || (s.k_has0 && r == 0)) ? 0 : s.k_base + (long long)(r - s.k_has0) * s.k_stride;
    cp_async16(sQ + swz(r, c), qb + qt * rs + c * 8, okq);
    cp_async16(sK + swz(r, c), qb + kt * rs + ro + c * 8, okk);
    cp_async16(sV + swz(r, c), qb + kt * rs + 2 * ro + c * 8, okk);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (warp * 16 >= s.nq) return;
  uint32_t qa[4][4];
  load_a_frags(sQ, warp * 16, lane, qa);
  float sc[2 * NW][4];
#pragma unroll
  for (int j = 0; j < 2 * NW; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int qtok0 = s.q_base + row0 * s.q_stride, qtok1 = s.q_base + row1 * s.q_stride;
  // causal: keys beyond this warp's last query are never needed
  const int k_lim = a.causal ? min(nk, warp * 16 + 16) : nk;
#pragma unroll
  for (int j = 0; j < 2 * NW; j += 2) {
    if (8 * j < k_lim) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t b0, b1, c0, c1;
        load_b_rows(sK, j, kk, lane, b0, b1, c0, c1);
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
  if (mx0 == -INFINITY) mx0 = 0.f;     // padded query rows of a causal tile
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
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NW; ++kk) {
    if (16 * kk < k_lim) {
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t b0, b1, c0, c1;
        load_b_cols(sV, kk, jd, lane, b0, b1, c0, c1);
        mma16816(o[jd], pa[kk], b0, b1);
        mma16816(o[jd + 1], pa[kk], c0, c1);
      }
    }
  }
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __syncwarp();     // this warp's Q rows are only read by this warp (fragments already in registers): reuse them for staging
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(smem + swz(row0, j) + 4 * t4) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(smem + swz(row1, j) + 4 * t4) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
  __syncwarp();
  bf16* ob = out + (long long)b * a.N * ro + (long long)h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 16 + i * 4 + (lane >> 3), c = lane & 7;
    if (r < s.nq)
      *reinterpret_cast<uint4*>(ob + (long long)(s.q_base + r * s.q_stride) * ro + c * 8) = *reinterpret_cast<const uint4*>(smem + swz(r, c));
  }
  if (t4 == 0) {
    float* lb = lse + ((long long)b * a.H + h) * a.N;
    if (row0 < s.nq) lb[qtok0] = mx0 * a.scale + __logf(l0);
    if (row1 < s.nq) lb[qtok1] = mx1 * a.scale + __logf(l1);
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 2) attn_group_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                     const float* __restrict__ lse, const float* __restrict__ delta,
                                                                     bf16* __restrict__ dqkv, AttnShape a) {
  constexpr int R = NW * 16;
  TVTS_DYN_SMEM(uint8_t, gsm, 128);                    // Q | K | V | dO | dQ-staging tiles, then lse*log2e [R], delta [R]
  const uint32_t sQ = smem_u32(gsm), sK = sQ + R * 128, sV = sK + R * 128, sD = sV + R * 128;
  uint8_t* gO = gsm + 4 * R * 128;
  float* Ls = reinterpret_cast<float*>(gsm + 5 * R * 128);
  float* Ds = Ls + R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const GroupSets s = group_sets(a, blockIdx.x);
  const int nk = s.k_has0 + s.nk_patch;
  const int nq_all = s.nq + s.k_has0;                 // space mode: the CLS query is row nq of the Q / dO tiles
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  bf16* dq_b = dqkv + (long long)b * a.N * rs + (long long)h * HD;
  auto q_token = [&](int r) -> long long { return r < s.nq ? s.q_base + (long long)r * s.q_stride : 0; };   // r == nq: CLS query
  for (int i = threadIdx.x; i < R * 8; i += NW * 32) {
    const int r = i >> 3, c = i & 7;
    const bool okq = r < nq_all, okk = r < nk;
    const long long qt = okq ? q_token(r) : 0;
    const long long kt = (!okk || (s.k_has0 && r == 0)) ? 0 : s.k_base + (long long)(r - s.k_has0) * s.k_stride;
    cp_async16(sQ + swz(r, c), qb + qt * rs + c * 8, okq);
    cp_async16(sD + swz(r, c), dob + qt * ro + c * 8, okq);
    cp_async16(sK + swz(r, c), qb + kt * rs + ro + c * 8, okk);
    cp_async16(sV + swz(r, c), qb + kt * rs + 2 * ro + c * 8, okk);
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
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
    const float l0 = Ls[row0], l1 = Ls[row1], d0 = Ds[row0], d1 = Ds[row1];
    const int k_lim = a.causal ? min(nk, warp * 16 + 16) : nk;
#pragma unroll 1
    for (int kc = 0; kc < k_lim; kc += 32) {
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t qa[4], da[4];       // A fragments are re-read per chunk (2 ldmatrix) instead of living in 32 registers
        ldsm_x4(sQ + swz(arow, 2 * kk + (lane >> 4)), qa[0], qa[1], qa[2], qa[3]);
        ldsm_x4(sD + swz(arow, 2 * kk + (lane >> 4)), da[0], da[1], da[2], da[3]);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (kc + 8 * j < k_lim) {
            uint32_t b0, b1, c0, c1;
            load_b_rows(sK, (kc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(sc[j], qa, b0, b1);
            mma16816(sc[j + 1], qa, c0, c1);
            load_b_rows(sV, (kc >> 3) + j, kk, lane, b0, b1, c0, c1);
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
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t b0, b1, c0, c1;
            load_b_cols(sK, (kc >> 4) + kk, jd, lane, b0, b1, c0, c1);
            mma16816(dq[jd], dsa[kk], b0, b1);
            mma16816(dq[jd + 1], dsa[kk], c0, c1);
          }
        }
      }
    }
    // dq of this warp's rows -> its private rows of the staging tile (frees 32 registers before pass B)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(gO + swz(row0, j) + 4 * t4) = pack_bf16x2(dq[j][0] * a.scale, dq[j][1] * a.scale);
      *reinterpret_cast<uint32_t*>(gO + swz(row1, j) + 4 * t4) = pack_bf16x2(dq[j][2] * a.scale, dq[j][3] * a.scale);
    }
  }
  // ---------------- pass B: rows = this warp's keys, columns = queries (+ CLS query) in chunks of 32: dK = scale * dS^T Q ; dV = P^T dO
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
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
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t ka[4], va[4];
        ldsm_x4(sK + swz(arow, 2 * kk + (lane >> 4)), ka[0], ka[1], ka[2], ka[3]);
        ldsm_x4(sV + swz(arow, 2 * kk + (lane >> 4)), va[0], va[1], va[2], va[3]);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (qc + 8 * j < nq_all) {
            uint32_t b0, b1, c0, c1;
            load_b_rows(sQ, (qc >> 3) + j, kk, lane, b0, b1, c0, c1);
            mma16816(sc[j], ka, b0, b1);
            mma16816(sc[j + 1], ka, c0, c1);
            load_b_rows(sD, (qc >> 3) + j, kk, lane, b0, b1, c0, c1);
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
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t b0, b1, c0, c1;
            load_b_cols(sQ, (qc >> 4) + kk, jd, lane, b0, b1, c0, c1);
            mma16816(dk[jd], dsa[kk], b0, b1);
            mma16816(dk[jd + 1], dsa[kk], c0, c1);
            load_b_cols(sD, (qc >> 4) + kk, jd, lane, b0, b1, c0, c1);
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
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(gsm + R * 128 + swz(row0, j) + 4 * t4) = pack_bf16x2(dk[j][0] * a.scale, dk[j][1] * a.scale);
    *reinterpret_cast<uint32_t*>(gsm + R * 128 + swz(row1, j) + 4 * t4) = pack_bf16x2(dk[j][2] * a.scale, dk[j][3] * a.scale);
    *reinterpret_cast<uint32_t*>(gsm + 2 * R * 128 + swz(row0, j) + 4 * t4) = pack_bf16x2(dv[j][0], dv[j][1]);
    *reinterpret_cast<uint32_t*>(gsm + 2 * R * 128 + swz(row1, j) + 4 * t4) = pack_bf16x2(dv[j][2], dv[j][3]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 16 + i * 4 + (lane >> 3), c = lane & 7;
    if (r < s.nq)
      *reinterpret_cast<uint4*>(dq_b + (s.q_base + (long long)r * s.q_stride) * rs + c * 8) = *reinterpret_cast<const uint4*>(gO + swz(r, c));
    if (r < nk && !(s.k_has0 && r == 0)) {            // key 0 = CLS in space mode: written by the cls_only launch
      const long long kt = s.k_base + (long long)(r - s.k_has0) * s.k_stride;
      *reinterpret_cast<uint4*>(dq_b + kt * rs + ro + c * 8) = *reinterpret_cast<const uint4*>(gsm + R * 128 + swz(r, c));
      *reinterpret_cast<uint4*>(dq_b + kt * rs + 2 * ro + c * 8) = *reinterpret_cast<const uint4*>(gsm + 2 * R * 128 + swz(r, c));
    }
  }
}

// number of warps a group-resident launch needs (0: the group does not fit -> streamed kernels)
inline int group_warps(const AttnShape& a) {
  if (a.mode == 2) return 0;
  const int nq = a.mode == 0 ? a.N : a.n + 1;         // backward appends the CLS query in space mode
  const int nk = a.mode == 0 ? a.N : a.n + 1;
  const int m = nq > nk ? nq : nk;
  const int nw = (m + 15) / 16;
  return (nw >= 2 && nw <= 7) ? nw : 0;
}

#ifndef TVTS_HOST_SHIM
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
#endif  // !TVTS_HOST_SHIM

// ================================================================================================ CLS row / column kernels
// In the divided modes the CLS token attends to (and is attended by) ALL N tokens: one query row and one key column per (batch,
// head) -- matrix-vector work.  One CTA per (batch, head): thread-per-token dot products over 16-byte row chunks (the vectors
// q0 / dO0 / k0 / v0 are broadcast from shared memory), block reductions for the softmax, then warp-split weighted row sums.
constexpr int CLS_THREADS = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float dot8(const uint4& u, const float* v) {
  float f[8];
  unpack8(u, f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(f[i], v[i], s);
  return s;
}
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {   // red: >= 9 floats; result broadcast to all
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = red[0];
    for (int w = 1; w < CLS_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    red[8] = r;
  }
  __syncthreads();
  return red[8];
}

// out[b, 0, h, :] = softmax(scale * q0 . K^T) V  over all N tokens; lse[b, h, 0]
__global__ void __launch_bounds__(CLS_THREADS) attn_cls_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                                                   AttnShape a) {
  TVTS_DYN_SMEM(float, csm, 16);                         // q0 [64] | red [16] | acc [8][64] | p [N]
  float* q0 = csm;
  float* red = csm + 64;
  float* accs = csm + 80;
  float* p = csm + 80 + 8 * 64;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + ro;
  const bf16* vb = kb + ro;
  // scores: 8 lanes per token (one 16-byte chunk each -> a warp reads 4 whole 128-byte rows per instruction), octet reduction
  const int sub = threadIdx.x & 7;
  float q0c[8];
  {
    const uint4 u = reinterpret_cast<const uint4*>(qb)[sub];
    unpack8(u, q0c);
#pragma unroll
    for (int i = 0; i < 8; ++i) q0c[i] *= a.scale;
  }
  (void)q0;
  float mx = -INFINITY;
  constexpr int TPI = CLS_THREADS / 8;                   // tokens per CTA iteration
  for (int j0 = 0; j0 < a.N; j0 += 4 * TPI) {            // 4 independent row loads in flight per thread
    uint4 kr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      kr[u] = j < a.N ? reinterpret_cast<const uint4*>(kb + (long long)j * rs)[sub] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      float sj = dot8(kr[u], q0c);
      sj += __shfl_xor_sync(0xffffffffu, sj, 1);
      sj += __shfl_xor_sync(0xffffffffu, sj, 2);
      sj += __shfl_xor_sync(0xffffffffu, sj, 4);
      if (j < a.N) {
        if (sub == 0) p[j] = sj;
        mx = fmaxf(mx, sj);
      }
    }
  }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < a.N; j += CLS_THREADS) {
    const float e = __expf(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);                    // (its barriers also publish p[])
  // weighted row sum: 8 lanes per token again (16-byte loads, 8 rows in flight per lane); lane (octet group g8, chunk sub)
  // accumulates dims 8*sub..8*sub+7 over tokens j = g8 (mod 32); groups are reduced by shuffles, warps through shared memory
  float acc8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc8[i] = 0.f;
  for (int j0 = 0; j0 < a.N; j0 += 8 * TPI) {
    uint4 vr[8];
    float pj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      const bool ok = j < a.N;
      vr[u] = ok ? reinterpret_cast<const uint4*>(vb + (long long)j * rs)[sub] : make_uint4(0, 0, 0, 0);
      pj[u] = ok ? p[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float f[8];
      unpack8(vr[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc8[i] = fmaf(pj[u], f[i], acc8[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc8[i] += __shfl_xor_sync(0xffffffffu, acc8[i], 8);
    acc8[i] += __shfl_xor_sync(0xffffffffu, acc8[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) accs[warp * 64 + 8 * lane + i] = acc8[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < CLS_THREADS / 32; ++w) o += accs[w * 64 + threadIdx.x];
    out[(long long)b * a.N * ro + (long long)h * HD + threadIdx.x] = opnd_from_float(o / sum);
  }
  if (threadIdx.x == 0) lse[((long long)b * a.H + h) * a.N] = mx + __logf(sum);
}

// dq of the CLS query (row 0 of the probability matrix) and dk / dv of the CLS key (column 0), all of token 0
__global__ void __launch_bounds__(CLS_THREADS) attn_cls_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                   const float* __restrict__ lse, const float* __restrict__ delta,
                                                                   bf16* __restrict__ dqkv, AttnShape a) {
  TVTS_DYN_SMEM(float, csm, 16);                         // q0 | do0 | k0 | v0 [64 each] | acc [8][192] | dsA [N] | pB [N] | dsB [N]
  float* q0 = csm;
  float* do0 = csm + 64;
  float* k0 = csm + 128;
  float* v0 = csm + 192;
  float* accs = csm + 256;
  float* dsA = csm + 256 + 8 * 192;
  float* pB = dsA + a.N;
  float* dsB = pB + a.N;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rs = 3LL * a.H * HD, ro = (long long)a.H * HD;
  const bf16* qb = qkv + (long long)b * a.N * rs + (long long)h * HD;
  const bf16* kb = qb + ro;
  const bf16* vb = kb + ro;
  const bf16* dob = dout + (long long)b * a.N * ro + (long long)h * HD;
  const float* lse_b = lse + ((long long)b * a.H + h) * a.N;
  const float* delta_b = delta + ((long long)b * a.H + h) * a.N;
  // pass 1: 8 lanes per token (16-byte chunks: a warp reads 4 whole rows per instruction), the four vectors live in registers
  const int sub = threadIdx.x & 7;
  float q0c[8], do0c[8], k0c[8], v0c[8];
  unpack8(reinterpret_cast<const uint4*>(qb)[sub], q0c);
  unpack8(reinterpret_cast<const uint4*>(dob)[sub], do0c);
  unpack8(reinterpret_cast<const uint4*>(kb)[sub], k0c);
  unpack8(reinterpret_cast<const uint4*>(vb)[sub], v0c);
#pragma unroll
  for (int i = 0; i < 8; ++i) { q0c[i] *= a.scale; k0c[i] *= a.scale; }   // s = (scale q) . k
  (void)q0; (void)do0; (void)k0; (void)v0;
  const float lse0 = lse_b[0], delta0 = delta_b[0];
  constexpr int TPI = CLS_THREADS / 8;
  for (int j0 = 0; j0 < a.N; j0 += 2 * TPI) {            // 8 independent 16-byte loads in flight per thread
    uint4 rk[2], rv[2], rq[2], rd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      const bool ok = j < a.N;
      const long long jj = ok ? j : 0;
      rk[u] = reinterpret_cast<const uint4*>(kb + jj * rs)[sub];
      rv[u] = reinterpret_cast<const uint4*>(vb + jj * rs)[sub];
      rq[u] = reinterpret_cast<const uint4*>(qb + jj * rs)[sub];
      rd[u] = reinterpret_cast<const uint4*>(dob + jj * ro)[sub];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      float sA = dot8(rk[u], q0c), dpA = dot8(rv[u], do0c), sB = dot8(rq[u], k0c), dpB = dot8(rd[u], v0c);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sA += __shfl_xor_sync(0xffffffffu, sA, o);
        dpA += __shfl_xor_sync(0xffffffffu, dpA, o);
        sB += __shfl_xor_sync(0xffffffffu, sB, o);
        dpB += __shfl_xor_sync(0xffffffffu, dpB, o);
      }
      if (j < a.N && sub == 0) {
        const float pa = __expf(sA - lse0);
        dsA[j] = pa * (dpA - delta0);
        const float pb = __expf(sB - lse_b[j]);
        pB[j] = pb;
        dsB[j] = pb * (dpB - delta_b[j]);
      }
    }
  }
  __syncthreads();
  // pass 2: weighted row sums in the same octet layout (12 independent 16-byte loads in flight per lane)
  float dq8[8], dk8[8], dv8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq8[i] = dk8[i] = dv8[i] = 0.f;
  for (int j0 = 0; j0 < a.N; j0 += 4 * TPI) {
    uint4 rk[4], rq[4], rd[4];
    float da[4], db[4], pb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * TPI + (threadIdx.x >> 3);
      const bool ok = j < a.N;
      const long long jj = ok ? j : 0;
      rk[u] = reinterpret_cast<const uint4*>(kb + jj * rs)[sub];
      rq[u] = reinterpret_cast<const uint4*>(qb + jj * rs)[sub];
      rd[u] = reinterpret_cast<const uint4*>(dob + jj * ro)[sub];
      da[u] = ok ? dsA[j] : 0.f;
      db[u] = ok ? dsB[j] : 0.f;
      pb[u] = ok ? pB[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fk[8], fq[8], fd[8];
      unpack8(rk[u], fk); unpack8(rq[u], fq); unpack8(rd[u], fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dq8[i] = fmaf(da[u], fk[i], dq8[i]);
        dk8[i] = fmaf(db[u], fq[i], dk8[i]);
        dv8[i] = fmaf(pb[u], fd[i], dv8[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dq8[i] += __shfl_xor_sync(0xffffffffu, dq8[i], 8); dq8[i] += __shfl_xor_sync(0xffffffffu, dq8[i], 16);
    dk8[i] += __shfl_xor_sync(0xffffffffu, dk8[i], 8); dk8[i] += __shfl_xor_sync(0xffffffffu, dk8[i], 16);
    dv8[i] += __shfl_xor_sync(0xffffffffu, dv8[i], 8); dv8[i] += __shfl_xor_sync(0xffffffffu, dv8[i], 16);
  }
  float* ar = accs + warp * 192;
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ar[8 * lane + i] = dq8[i];
      ar[64 + 8 * lane + i] = dk8[i];
      ar[128 + 8 * lane + i] = dv8[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < 192) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < CLS_THREADS / 32; ++w) o += accs[w * 192 + threadIdx.x];
    const int which = threadIdx.x >> 6, dim = threadIdx.x & 63;       // 0 dq, 1 dk, 2 dv
    if (which < 2) o *= a.scale;
    dqkv[(long long)b * a.N * rs + (long long)which * ro + (long long)h * HD + dim] = opnd_from_float(o);
  }
}

// The CLS row / column of the divided modes (one query over all N keys, one key under all N queries: 384 tiny CTAs that stream
// 13 tiles each) is latency-bound and independent of the group kernels (disjoint outputs), so it runs on a side stream forked
// from / joined back into the caller's stream with events -- also legal inside a CUDA-graph capture, where it becomes a parallel
// branch of the graph.
#ifndef TVTS_HOST_SHIM
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}
int g_attn_side = 1;
#endif  // !TVTS_HOST_SHIM

inline bool use_time_kernels(const AttnShape& a) { return a.mode == 2 && a.T <= T_MAX; }

}  // namespace

#ifndef TVTS_HOST_SHIM
extern "C" int tvts_attn_set_side_stream(int on) {
  g_attn_side = on;
  return TVTS_OK;
}

extern "C" int tvts_attn_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T,
                             int64_t n, int64_t causal, float scale, void* stream) {
  if (d != HD) return tvts_attn_generic_fwd(qkv, out, lse, B, N, H, d, mode, T, n, causal, scale, stream);   // attention_hd.cu
  if (tvts_attn_tc_supported(B, N, H, d, mode, T, n, causal))                                                 // attention_tc.cu (tcgen05)
    return tvts_attn_tc_fwd(qkv, out, lse, B, N, H, d, mode, T, n, causal, scale, stream);
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && lse, "attn_fwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gw = group_warps(a);
  const bool time_k = use_time_kernels(a);
  const bool split = a.mode != 0 && (time_k || gw);       // group/time kernel + separate CLS launch
  SideStream* sd = (split && g_attn_side) ? side_stream() : nullptr;
  cudaStream_t cls_st = st;
  if (sd) {
    TVTS_CHECK_CUDA(cudaEventRecord(sd->fork, st));
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
    cls_st = sd->stream;
  }
  if (split) {                                             // CLS query over all tokens
    const int smem_bytes = (80 + 8 * 64 + a.N) * 4;
    if (smem_bytes <= 48 * 1024) {
      attn_cls_fwd_kernel<<<dim3((unsigned)H, (unsigned)B), CLS_THREADS, smem_bytes, cls_st>>>((const bf16*)qkv, (bf16*)out, lse, a);
    } else {
      AttnShape c = a;
      c.cls_only = 1;
      dim3 grid(num_blocks_x(c), (unsigned)H, (unsigned)B);
      attn_fwd_kernel<<<grid, kThreads, 0, cls_st>>>((const bf16*)qkv, (bf16*)out, lse, c);
    }
    TVTS_LAUNCH_CHECK();
  }
  if (time_k) {
    dim3 tg((unsigned)((a.n + TW - 1) / TW), (unsigned)H, (unsigned)B);
    attn_time_fwd_kernel<<<tg, TW * 32, 0, st>>>((const bf16*)qkv, (bf16*)out, lse, a);
    TVTS_LAUNCH_CHECK();
  } else if (gw) {
    dim3 gg((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)H, (unsigned)B);
    switch (gw) {
#define TVTS_GF(NW_) case NW_: attn_group_fwd_kernel<NW_><<<gg, NW_ * 32, 0, st>>>((const bf16*)qkv, (bf16*)out, lse, a); break;
      TVTS_GF(2) TVTS_GF(3) TVTS_GF(4) TVTS_GF(5) TVTS_GF(6) TVTS_GF(7)
#undef TVTS_GF
    }
    TVTS_LAUNCH_CHECK();
  } else {
    dim3 grid(num_blocks_x(a), (unsigned)H, (unsigned)B);
    attn_fwd_kernel<<<grid, kThreads, 0, st>>>((const bf16*)qkv, (bf16*)out, lse, a);
    TVTS_LAUNCH_CHECK();
  }
  if (sd) {
    TVTS_CHECK_CUDA(cudaEventRecord(sd->join, sd->stream));
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
  }
  return TVTS_OK;
}

extern "C" int tvts_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int64_t B,
                             int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale,
                             void* stream) {
  if (d != HD)                                                                                               // attention_hd.cu
    return tvts_attn_generic_bwd(qkv, out, dout, lse, delta_ws, dqkv, B, N, H, d, mode, T, n, causal, scale, stream);
  if (tvts_attn_tc_supported(B, N, H, d, mode, T, n, causal))                                                 // attention_tc.cu (tcgen05)
    return tvts_attn_tc_bwd(qkv, out, dout, lse, dqkv, B, N, H, d, mode, T, n, causal, scale, stream);
  AttnShape a{(int)B, (int)N, (int)H, (int)mode, (int)T, (int)n, (int)causal, scale, 0, 0, 0};
  if (B == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv, "attn_bwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)B * N * H;
  attn_delta_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta_ws, (int)B, (int)N, (int)H);
  TVTS_LAUNCH_CHECK();
  const int gw = group_warps(a);
  const bool time_k = use_time_kernels(a);
  const bool split = a.mode != 0 && (time_k || gw);
  SideStream* sd = (split && g_attn_side) ? side_stream() : nullptr;
  cudaStream_t cls_st = st;
  if (sd) {
    TVTS_CHECK_CUDA(cudaEventRecord(sd->fork, st));      // after delta
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
    cls_st = sd->stream;
  }
  if (split) {                                             // dq of the CLS query, dk/dv of the CLS key
    const int smem_bytes = (256 + 8 * 192 + 3 * a.N) * 4;
    if (smem_bytes <= 48 * 1024) {
      attn_cls_bwd_kernel<<<dim3((unsigned)H, (unsigned)B), CLS_THREADS, smem_bytes, cls_st>>>((const bf16*)qkv, (const bf16*)dout, lse,
                                                                                                delta_ws, (bf16*)dqkv, a);
      TVTS_LAUNCH_CHECK();
    } else {
      AttnShape c = a;
      c.cls_only = 1;
      dim3 grid(num_blocks_x(c), (unsigned)H, (unsigned)B);
      attn_bwd_kernel<0><<<grid, kThreads, 0, cls_st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, c);
      TVTS_LAUNCH_CHECK();
      attn_bwd_kernel<1><<<grid, kThreads, 0, cls_st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, c);
      TVTS_LAUNCH_CHECK();
    }
  }
  if (time_k) {
    dim3 tg((unsigned)((a.n + TW - 1) / TW), (unsigned)H, (unsigned)B);
    attn_time_bwd_kernel<<<tg, TW * 32, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
    TVTS_LAUNCH_CHECK();
  } else if (gw) {
    dim3 gg((unsigned)(a.mode == 0 ? 1 : a.T), (unsigned)H, (unsigned)B);
    const int smem_bytes = 5 * gw * 16 * 128 + 2 * gw * 16 * 4;
    switch (gw) {
#define TVTS_GB(NW_)                                                                                                              \
  case NW_: {                                                                                                                     \
    static bool set_##NW_ = false;                                                                                                \
    if (!set_##NW_) {                                                                                                             \
      TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_group_bwd_kernel<NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
      set_##NW_ = true;                                                                                                           \
    }                                                                                                                             \
    attn_group_bwd_kernel<NW_><<<gg, NW_ * 32, smem_bytes, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a); \
  } break;
      TVTS_GB(2) TVTS_GB(3) TVTS_GB(4) TVTS_GB(5) TVTS_GB(6) TVTS_GB(7)
#undef TVTS_GB
    }
    TVTS_LAUNCH_CHECK();
  } else {
    dim3 grid(num_blocks_x(a), (unsigned)H, (unsigned)B);
    attn_bwd_kernel<0><<<grid, kThreads, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
    TVTS_LAUNCH_CHECK();
    attn_bwd_kernel<1><<<grid, kThreads, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
    TVTS_LAUNCH_CHECK();
  }
  if (sd) {
    TVTS_CHECK_CUDA(cudaEventRecord(sd->join, sd->stream));
    TVTS_CHECK_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
  }
  return TVTS_OK;
}

// tvts_attn_bwd + the bias gradient of the qkv Linear (dbias[3*H*d] += column sums of dqkv over all B*N tokens): fused into the tcgen05
// kernels' epilogue where those run, otherwise the plain backward followed by the column-sum kernel
extern "C" int tvts_attn_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                  float* dbias, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n,
                                  int64_t causal, float scale, void* stream) {
  TVTS_REQUIRE(dbias != nullptr, "attn_bwd_bias: null dbias");
  if (d == HD && tvts_attn_tc_supported(B, N, H, d, mode, T, n, causal))
    return tvts_attn_tc_bwd_bias(qkv, out, dout, lse, dqkv, dbias, B, N, H, d, mode, T, n, causal, scale, stream);
  int rc = tvts_attn_bwd(qkv, out, dout, lse, delta_ws, dqkv, B, N, H, d, mode, T, n, causal, scale, stream);
  if (rc) return rc;
  return tvts_colsum_bf16(dqkv, dbias, B * N, 3 * H * d, 3 * H * d, stream);
}

// ------------------------------------------------------------------------------------------------ query-window attention
// Full (non-causal) attention where only rows [q0, q0+qn) of every sample act as QUERIES while all N tokens are keys / values:
// the last block of the sort head feeds the loss through its n_trans transcript rows only (v2/model/sort_transformer.py:134-142),
// so its attention, projection and MLP are evaluated for those rows alone (identical results, ~S/n_trans times less work).
extern "C" int tvts_attn_window_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t q0,
                                    int64_t qn, float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, 0, 0, 0, 0, scale, 0, (int)q0, (int)qn};
  if (B == 0 || qn == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && lse && q0 >= 0 && qn > 0 && q0 + qn <= N, "attn_window_fwd: bad arguments");
  dim3 grid(num_blocks_x(a), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)qkv, (bf16*)out, lse, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

// dq is written for the window rows only (the caller zero-fills dqkv first); dk / dv are written for every token.
// `out` / `dout` rows outside the window are never read by the dQ / dK,dV kernels (delta_ws is filled for all rows).
extern "C" int tvts_attn_window_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                    int64_t B, int64_t N, int64_t H, int64_t d, int64_t q0, int64_t qn, float scale, void* stream) {
  AttnShape a{(int)B, (int)N, (int)H, 0, 0, 0, 0, scale, 0, (int)q0, (int)qn};
  if (B == 0 || qn == 0) return TVTS_OK;
  int rc = check_shape(a, d);
  if (rc) return rc;
  TVTS_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv && q0 >= 0 && qn > 0 && q0 + qn <= N, "attn_window_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)B * N * H;
  attn_delta_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta_ws, (int)B, (int)N, (int)H);
  TVTS_LAUNCH_CHECK();
  dim3 g0(num_blocks_x(a, false), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<0><<<g0, kThreads, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  dim3 g1(num_blocks_x(a, true), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<1><<<g1, kThreads, 0, st>>>((const bf16*)qkv, (const bf16*)dout, lse, delta_ws, (bf16*)dqkv, a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
#endif  // !TVTS_HOST_SHIM
