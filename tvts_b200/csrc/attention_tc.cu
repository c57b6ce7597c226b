// tcgen05 attention for the SHORT groups of the hot path (head dim 64): the divided-space attention of the video tower
// (v2/model/video_encoder_ViT_B_16.py:38-76, `'b (f n) d -> (b f) n d'`: n kept patches of one frame + the CLS key, 99 rows for ViT-B/16
// at mask 0.5) and one-tile full / causal sequences (the 77-token CLIP text tower, v2/CLIP/clip/model.py:171-203).
//
// One TILE = one (batch, head, group).  The whole group fits ONE 128-row UMMA tile, so there is no streaming loop and no online
// rescaling; persistent CTAs (one per resident slot) walk the tiles and overlap the next tile's loads with the current tile's tail:
//   forward : TMA(Q,K,V) -> S = Q K^T (tcgen05.mma, fp32 in TMEM) -> row softmax straight out of TMEM (tcgen05.ld: one thread = one
//             query row, no shuffles) -> P (16-bit, 128B-swizzled shared memory) -> O = P V (tcgen05.mma) -> TMA store
//   backward: TMA(Q,K,V,dO) -> S, dP = dO V^T -> P = exp(S - lse), delta = rowsum(P o dP) (in the tile: no separate delta pass) ->
//             dV = P^T dO -> dS = P o (dP - delta) * scale -> dK = dS^T Q, dQ = dS K -> three TMA stores
// Every tile is read from HBM exactly once and every output written once: the kernels are HBM-bound by construction
// (forward 4 x n x 128 B per tile, backward 7 x n x 128 B).
//
// CLS token (space mode).  The CLS QUERY attends to every token of the clip and the CLS KEY is seen by every query.  It rides along as
// row / column n of every frame tile: the tile computes the CLS query's softmax over ITS keys (the CLS key itself is counted by frame
// 0 only) and writes a normalised partial (o, lse) to a small workspace; per output column, the thread that draws the LAST ticket of
// that column merges the partials (log-sum-exp combine, fixed order).  The backward does the same with partial dq(CLS query), dk / dv
// (CLS key) sums.
// This replaces the separate latency-bound attn_cls_* launches (and their side stream) of the mma.sync path; there is no second launch.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

constexpr int HD = 64;
constexpr uint32_t ATOM = 16384;   // one [128 rows x 128 B] 128B-swizzled tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t kOperandFormat = TVTS_OPERAND_IS_FP16 ? 0u : ((1u << 7) | (1u << 10));
constexpr size_t kTicketBytes = 8u << 20;      // per-device ticket area of the CLS merge: 192 int tickets per (b, h)
constexpr int kTcThreads = 256;    // 8 warps: TMEM lane quadrant = warp % 4, column half = warp / 4 (two threads share one row)

// -DTVTS_ATTN_PROF (measurement builds only, tools/attn_phase_prof.py): thread 0 of every CTA stamps clock64 at the phase boundaries
#ifdef TVTS_ATTN_PROF
constexpr int kProfStamps = 12, kProfTiles = 8192;
__device__ long long g_prof[kProfTiles][kProfStamps + 1];
#define PROF_STAMP(i)                                                                                   \
  do {                                                                                                  \
    if (threadIdx.x == 0 && t < kProfTiles) {                                                           \
      g_prof[t][i] = clock64();                                                                         \
      if ((i) == 0) {                                                                                   \
        unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); g_prof[t][kProfStamps] = sm_;    \
        unsigned long long ns_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_));                 \
        g_prof[t][10] = (long long)ns_; g_prof[t][11] = blockIdx.x;                                     \
      }                                                                                                 \
    }                                                                                                   \
  } while (0)
#define PROF_END()                                                                                      \
  do {                                                                                                  \
    if (threadIdx.x == 0 && t < kProfTiles) {                                                           \
      unsigned long long ns_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_)); g_prof[t][9] = (long long)ns_; \
    }                                                                                                   \
  } while (0)
#else
#define PROF_STAMP(i) do {} while (0)
#define PROF_END() do {} while (0)
#endif

struct TcShape {
  int B, N, H;
  int mode;     // 0: one tile per (b, h) = the whole sequence (N <= 128); 1: space (tile = frame); 2: time (tile = GP patch positions x T frames)
  int T, n;
  int causal;   // mode 0 only
  float scale;
  int GP;       // mode 2: patch positions per tile (row r of the tile = frame r / GP, position g * GP + r % GP)
  int chunks;   // tiles per (b, h): 1 | T | ceil(n / GP)
  int rows;     // rows the TMA box moves per tile: N | n | T * GP
  int L;        // rows = keys of the tile: `rows` (+ 1 in modes 1 / 2: the CLS token, last)
  int LP;       // L rounded up to a multiple of 16 (UMMA N / K granularity)
  int ahead;    // backward (persistent): 1 = L2-prefetch the next tile's boxes and plain loads one tile early
  int ahead_tiles;   // forward (one CTA per tile): L2 prefetch distance in tiles (= CTAs resident on the chip), 0 = off
};

__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// tcgen05.ld is .sync.aligned: reconverge the warp first (per-row predicates / the single CLS-row thread may have diverged it)
__device__ __forceinline__ void tmem_ld_row32(uint32_t taddr, uint32_t* v) {
  __syncwarp();
  tmem_ld_32x32(taddr, v);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 32 bytes (8 packed registers) to global memory in one instruction (STG.256; sm_100+); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(src_smem),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// instruction descriptor (kind::f16): D = f32, A/B = the library's 16-bit operand format, majors at bits 15 / 16 (1 = MN-major), N >> 3 at 17,
// M >> 4 at 24 (same encoding as gemm_tcgen05.cu)
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | kOperandFormat | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// K-major operand tile(s) [rows x 64 k] per 16 KB atom: k-step ks (16 elements) lives in atom ks / 4 at byte offset (ks % 4) * 32
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int ks) {
  return umma_smem_desc(tile + (uint32_t)(ks >> 2) * ATOM + (uint32_t)(ks & 3) * 32u, 16, 1024);
}
// MN-major operand: tile(s) [k rows x 64 mn] -- 16 k-rows per step = 2048 B; 64-wide MN groups are one atom apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int ks) { return umma_smem_desc(tile + (uint32_t)ks * 2048u, ATOM, 1024); }

// L2 prefetch of plain global memory (the per-row statistics and the CLS rows the NEXT tile of this SM slot will read with ld.global)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// lanes [lane0, lane0 + nl) of a warp touch every 128-byte line of [p, p + bytes)
__device__ __forceinline__ void prefetch_l2_range(const void* p, int bytes, int idx, int nl) {
  const uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127, hi = reinterpret_cast<uintptr_t>(p) + (uintptr_t)bytes;
  for (uintptr_t x = lo + (uintptr_t)idx * 128u; x < hi; x += (uintptr_t)nl * 128u) prefetch_l2(reinterpret_cast<const void*>(x));
}

// ---- tile geometry ---------------------------------------------------------------------------------------------------------------
// token (within the sample) of tile row r < rows, or -1 when the row is padding (mode 2: positions past n in the last tile)
__device__ __forceinline__ int row_token(const TcShape& a, int g, int r) {
  if (a.mode == 0) return r;
  if (a.mode == 1) return 1 + g * a.n + r;
  const int t = r / a.GP, i = r - t * a.GP, pos = g * a.GP + i;
  return pos < a.n ? 1 + t * a.n + pos : -1;
}
// bit j = key column (col0 + j) is attended by query row r (64 columns: this thread's half of the row)
__device__ __forceinline__ unsigned long long key_mask(const TcShape& a, int g, int r, int col0) {
  const bool cls_q = a.mode != 0 && r == a.rows;
  unsigned long long m = 0ull;
  if (r >= a.L) return 0ull;
  if (a.mode != 2) {
    int kend = a.L;
    if (a.causal) kend = min(kend, r + 1);
    if (cls_q && g != 0) kend = a.rows;                   // CLS query: the CLS key (last column) is counted by tile 0 only
    const int k = min(max(kend - col0, 0), 64);
    return k == 64 ? ~0ull : ((1ull << k) - 1ull);
  }
  const int nvalid = min(a.GP, a.n - g * a.GP);           // positions of this tile that exist
  if (!cls_q) {
    const int t = r / a.GP, i = r - t * a.GP;
    if (i >= nvalid) return 0ull;
    for (int tt = 0; tt < a.T; ++tt) {                    // the same position in every frame
      const int j = tt * a.GP + i - col0;
      if (j >= 0 && j < 64) m |= 1ull << j;
    }
    const int jc = a.rows - col0;                         // + the CLS key
    if (jc >= 0 && jc < 64) m |= 1ull << jc;
    return m;
  }
  for (int tt = 0; tt < a.T; ++tt) {                      // CLS query: every existing patch key of the tile
    const int lo = max(tt * a.GP - col0, 0), hi = min(tt * a.GP + nvalid - col0, 64);
    if (hi > lo) m |= (hi - lo == 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull)) << lo;
  }
  const int jc = a.rows - col0;
  if (g == 0 && jc >= 0 && jc < 64) m |= 1ull << jc;
  return m;
}
// TMA coordinates of tile g of sample b (4-D maps: {column, position / row, frame, sample})
__device__ __forceinline__ void tile_coords(const TcShape& a, int g, int& c1, int& c2) {
  c1 = a.mode == 2 ? g * a.GP : 0;
  c2 = a.mode == 1 ? g : 0;
}

// ================================================================================================ persistent tile loop (backward)
// The BACKWARD kernel is persistent: the grid is one CTA per resident slot (2 per SM) and each CTA draws tiles (linear tile index t =
// (b * H + h) * chunks + g) from a dynamic scheduler, two tiles ahead.  Measured before this (round 2, profiles/r2/call15_*): of the
// 18 200 cycles a backward tile cost its slot, 2 300 were the gap between two CTAs, 4 200 the per-CTA prologue (TMEM allocation, barrier
// init, the latency of the plain loads of lse / CLS rows) and 3 600 the tail (CLS publish + fence + ticket round trip).  In the loop
//   * TMEM, barriers, tensor-map prefetch and the zero rows happen once per CTA;
//   * the outputs leave straight from registers (32-byte stores), so every operand buffer is free for the NEXT tile's box as soon as
//     its last MMA is done (V behind bar_1, dO behind pass B, Q / K behind bar_2) and no store read-out gates a load;
//   * the next tile's boxes are L2-prefetched one tile early, its plain loads (CLS rows, log-sum-exp, masks) are taken in the tail of
//     the current tile;
//   * the CLS partial of tile i is published at the top of tile i + 1 and its ticket drawn in the middle of it, where the warps wait
//     for an MMA anyway.
// Result (c3 shape, cold L2, CUDA events, profiles/r2/call23_attn_bench.txt): 117 -> 110.5 us (space), 116 -> 115 us (time).  What bounds
// it now is the dependent chain of a tile (three MMA round trips of ~1 000-1 500 cycles each between the two passes and the epilogue)
// with only two tiles in flight per SM (98 KB of shared memory each): one loop iteration stays at ~15 600 cycles however the work is
// redistributed between its phases (profiles/r2/call17-23_attn_phase_*).  The FORWARD kernel stays one CTA per tile (four resident per
// SM, 51 KB each): the same persistent structure measured slower there (55 vs 51 us: three CTAs per SM once the loop state pushes it
// past 64 registers; capped at 64 it spills).
struct TileId { int b, h, g, c1, c2; };
__device__ __forceinline__ TileId decode_tile(const TcShape& a, unsigned t) {      // B * H * chunks < 2^31 (checked by the host side)
  TileId x;
  const unsigned bh = t / (unsigned)a.chunks;
  x.g = (int)(t - bh * (unsigned)a.chunks);
  x.b = (int)(bh / (unsigned)a.H);
  x.h = (int)(bh - (unsigned)x.b * (unsigned)a.H);
  tile_coords(a, x.g, x.c1, x.c2);
  return x;
}

// forward: the tile `ahead_tiles` further on in linear block order (CTAs are dispatched in that order: with ahead_tiles = the CTAs
// resident on the chip it is the tile the same SM slot runs one wave from now)
__device__ __forceinline__ bool next_tile(const TcShape& a, int b, int h, int g, int& b2, int& h2, int& g2) {
  const long long nl = ((long long)b * a.H + h) * a.chunks + g + a.ahead_tiles;
  if (a.ahead_tiles <= 0 || nl >= (long long)a.B * a.H * a.chunks) return false;
  g2 = (int)(nl % a.chunks);
  h2 = (int)((nl / a.chunks) % a.H);
  b2 = (int)(nl / ((long long)a.chunks * a.H));
  return true;
}

// ================================================================================================ forward
// shared memory: Q | K | V tiles (P overlays Q|K once S is complete; the output staging tile overlays P), row-statistics exchange
// between the two column halves, 3 mbarriers, TMEM holder
constexpr int FWD_XCH = 3 * ATOM;                       // float [2][2][128]
constexpr int FWD_BAR = FWD_XCH + 2048;
constexpr int FWD_SMEM = FWD_BAR + 64 + 1024;

__global__ void __launch_bounds__(kTcThreads, 3)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, const bf16* __restrict__ qkv,
                   bf16* __restrict__ out, float* __restrict__ lse, float* __restrict__ cls_ws, int* __restrict__ tickets, TcShape a) {
  __shared__ float s_cls[65];     // the CLS query's normalised partial output of this tile (64 columns) and its log-sum-exp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = base + ATOM, sV = base + 2 * ATOM, sP = base;
  float* xch = reinterpret_cast<float*>(gen + FWD_XCH);
  const uint32_t bar_ld = base + FWD_BAR, bar_s = bar_ld + 8, bar_o = bar_ld + 16, holder = bar_ld + 24;
  volatile uint32_t* holder_gen = reinterpret_cast<volatile uint32_t*>(gen + FWD_BAR + 24);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const long long ld = 3LL * a.H * HD;
  int c1, c2;
  tile_coords(a, g, c1, c2);

  if (warp == 0) {
    tmem_alloc(holder, 128);
  } else if (warp == 1) {
    // pull the tile that will run in this slot one wave from now into L2 (CTAs are dispatched in linear block order): its TMA
    // loads -- and the plain loads of its CLS rows -- then see L2 latency instead of HBM latency: the kernel is latency-bound per
    // tile, not bandwidth-bound
    int b2 = 0, h2 = 0, g2 = 0;
    const bool nxt = next_tile(a, b, h, g, b2, h2, g2);
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_out);
      mbar_init(bar_ld, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(bar_ld, 3u * (uint32_t)a.rows * 128u);
      tma_load_4d(sQ, &tm_qkv, bar_ld, h * HD, c1, c2, b);
      tma_load_4d(sK, &tm_qkv, bar_ld, a.H * HD + h * HD, c1, c2, b);
      tma_load_4d(sV, &tm_qkv, bar_ld, 2 * a.H * HD + h * HD, c1, c2, b);
      if (nxt) {
        int d1, d2;
        tile_coords(a, g2, d1, d2);
        tma_prefetch_4d(&tm_qkv, h2 * HD, d1, d2, b2);
        tma_prefetch_4d(&tm_qkv, a.H * HD + h2 * HD, d1, d2, b2);
        tma_prefetch_4d(&tm_qkv, 2 * a.H * HD + h2 * HD, d1, d2, b2);
      }
    } else if (nxt && a.mode != 0 && lane <= 3) {      // the CLS token's q / k / v rows of that tile's sample (one 128-byte line each)
      prefetch_l2(qkv + (long long)b2 * a.N * ld + (long long)(lane - 1) * a.H * HD + h2 * HD);
    }
  } else if (warp == 2) {
    if (a.mode != 0 && lane < 24) {      // the CLS token's q / k / v rows -> row `rows` of the three tiles (generic-proxy stores, swizzled by hand)
      const int m = lane >> 3, c = lane & 7;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + (long long)b * a.N * ld + (long long)m * a.H * HD + h * HD + c * 8));
      st_shared_v4(base + (uint32_t)m * ATOM + swz(a.rows, c), v.x, v.y, v.z, v.w);
    }
  } else if (warp == 3) {
    for (int i = lane; i < (a.LP - a.L) * 8; i += 32)      // V rows [L, LP) take part in P V with P = 0: they must be finite
      st_shared_v4(sV + swz(a.L + (i >> 3), i & 7), 0u, 0u, 0u, 0u);
  }
  const int q4 = warp & 3, half = warp >> 2;
  const int r = q4 * 32 + lane;                            // query row = TMEM lane
  const unsigned long long kmask = key_mask(a, g, r, half * 64);
  const uint32_t km[2] = {(uint32_t)kmask, (uint32_t)(kmask >> 32)};
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;

  if (warp == 1 && lane == 0) {
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(a.LP, false, false);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc, k > 0 ? 1u : 0u);
    umma_commit(bar_s);
  }
  __syncwarp();
  mbar_wait(bar_s, 0);
  tc_fence_after();

  // ---- softmax: two threads per query row (64 key columns each); S comes straight out of TMEM
  const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)half * 64u;
  const float sl2 = a.scale * LOG2E;
  uint32_t v[32];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (half * 64 + c * 32 < a.LP) {                       // warp-uniform
      tmem_ld_row32(trow + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((km[c] >> j) & 1u) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
  }
  xch[half * 128 + r] = mx;
  __syncthreads();
  mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
  if (!(fabsf(mx) < INFINITY)) mx = 0.f;                  // rows without any key (padding)
  float l = 0.f;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (half * 64 + c * 32 < a.LP) {
      tmem_ld_row32(trow + c * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ((km[c] >> (2 * j)) & 1u) ? exp2f((__uint_as_float(v[2 * j]) - mx) * sl2) : 0.f;
        const float p1 = ((km[c] >> (2 * j + 1)) & 1u) ? exp2f((__uint_as_float(v[2 * j + 1]) - mx) * sl2) : 0.f;
        l += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      const uint32_t atom = sP + (uint32_t)half * ATOM;
#pragma unroll
      for (int q = 0; q < 4; ++q) st_shared_v4(atom + swz(r, c * 4 + q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
  }
  xch[256 + half * 128 + r] = l;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();          // every row of P is in shared memory and every thread is done reading S (O overwrites its columns)

  if (warp == 1 && lane == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(HD, false, true);
    const int ksteps = a.LP >> 4;
    for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tmem, desc_kmajor(sP, ks), desc_mnmajor(sV, ks), idesc, ks > 0 ? 1u : 0u);
    umma_commit(bar_o);
  }
  __syncwarp();
  l += xch[256 + (half ^ 1) * 128 + r];
  mbar_wait(bar_o, 0);
  tc_fence_after();

  // ---- epilogue: O / l -> 16-bit rows in the staging tile (overlays P: the P V MMAs have completed) -> one TMA store
  const float inv = 1.0f / l;
  const bool cls_row = a.mode != 0 && r == a.rows;
  {
    tmem_ld_row32(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)half * 32u, v);     // this thread's 32 of the row's 64 output columns
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
#pragma unroll
    for (int q = 0; q < 4; ++q) st_shared_v4(sP + swz(r, half * 4 + q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    if (cls_row) {              // handed to 64 threads through shared memory: a single thread writing (and fencing) 65 words to global
#pragma unroll                  // memory used to hold the whole CTA at the barrier below
      for (int j = 0; j < 32; ++j) s_cls[half * 32 + j] = __uint_as_float(v[j]) * inv;
      if (half == 0) s_cls[64] = mx * a.scale + __logf(l);
    }
  }
  if (half == 0 && r < a.rows) {
    const int tok = row_token(a, g, r);
    if (tok >= 0) lse[((long long)b * a.H + h) * a.N + tok] = mx * a.scale + __logf(l);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tma_store_4d(&tm_out, sP, h * HD, c1, c2, b);
    bulk_commit();
  }
  if (a.mode != 0 && tid < 64) {
    // CLS query: thread j publishes column j of this tile's partial (o_j, lse) and takes a ticket of column j; the thread that draws the
    // LAST ticket of its column (any tile of this (b, h)) merges the column's partials (log-sum-exp combine, fixed order) ->
    // out[b, 0, h, j], lse[b, h, 0].  Per-column tickets: no CTA-wide barrier waits for an atomic's round trip.
    const long long bh = (long long)b * a.H + h;
    float2* ws2 = reinterpret_cast<float2*>(cls_ws) + bh * a.chunks * 64;
    ws2[g * 64 + tid] = make_float2(s_cls[tid], s_cls[64]);
    __threadfence();                               // the partial is visible before the ticket is
    int* tk = tickets + bh * 64 + tid;
    if (atomicAdd(tk, 1) == a.chunks - 1) {
      *tk = 0;                                     // ready for the next launch (stream-ordered)
      __threadfence();
      const volatile float* w = reinterpret_cast<const volatile float*>(ws2);
      float m = -INFINITY;
      for (int gg = 0; gg < a.chunks; ++gg) m = fmaxf(m, w[(gg * 64 + tid) * 2 + 1]);
      float acc = 0.f, sw = 0.f;
      for (int gg = 0; gg < a.chunks; ++gg) {
        const float e = __expf(w[(gg * 64 + tid) * 2 + 1] - m);
        acc = fmaf(e, w[(gg * 64 + tid) * 2], acc);
        sw += e;
      }
      out[(long long)b * a.N * a.H * HD + h * HD + tid] = opnd_from_float(acc / sw);
      if (tid == 0) lse[bh * a.N] = m + __logf(sw);
    }
  }
  if (tid == 0) bulk_wait_read<0>();     // the staging tile has been read out: the CTA may retire while the write drains
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ================================================================================================ backward
// shared memory: Q | K | V | dO tiles, a two-atom P / dS buffer, the delta exchange, 4 mbarriers, TMEM holder.  TMEM (256 columns):
// S [0,128) and dP [128,256); dV reuses [0,64), dK [64,128), dQ [128,192) once their previous contents have been consumed.  dV / dQ / dK
// leave straight from registers (two 32-byte stores per thread and tensor), so every operand buffer is free for the next tile's box as
// soon as its last MMA is done: V behind bar_1, dO behind bar_dv, Q and K behind bar_2.
constexpr int BWD_XCH = 6 * ATOM;                        // float [2][128]
constexpr int BWD_BAR = BWD_XCH + 1024;
constexpr int BWD_SMEM = BWD_BAR + 64 + 1024;

__global__ void __launch_bounds__(kTcThreads, 2)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const bf16* __restrict__ qkv,
                   const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ cls_ws,
                   float* __restrict__ dbias, bf16* __restrict__ dqkv, int* __restrict__ tickets, int* __restrict__ sched, TcShape a) {
  __shared__ float s_cls[192];    // this tile's partial dq (CLS query), dk, dv (CLS key), fp32
  __shared__ unsigned s_next;     // tile scheduler: the tile index drawn for the iteration after next
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = base + ATOM, sV = base + 2 * ATOM, sDO = base + 3 * ATOM, sP = base + 4 * ATOM;
  float* xch = reinterpret_cast<float*>(gen + BWD_XCH);
  const uint32_t bar_ld = base + BWD_BAR, bar_1 = bar_ld + 8, bar_dv = bar_ld + 16, bar_2 = bar_ld + 24, holder = bar_ld + 32;
  volatile uint32_t* holder_gen = reinterpret_cast<volatile uint32_t*>(gen + BWD_BAR + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 3LL * a.H * HD, ldo = (long long)a.H * HD;
  const unsigned total = (unsigned)(a.B * a.H * a.chunks), G = gridDim.x;
  const int q4 = warp & 3, half = warp >> 2;
  const int r = q4 * 32 + lane;
  const bool cls_row = a.mode != 0 && r == a.rows;
  const bool tma_thread = warp == 1 && lane == 0;
  const int ksteps = a.LP >> 4;
  const float sl2 = a.scale * LOG2E;
  const uint32_t box_bytes = (uint32_t)a.rows * 128u;

  unsigned t = blockIdx.x;
  TileId x = decode_tile(a, t);
  if (tid == 64) s_next = G + (unsigned)atomicAdd(sched, 1);      // this CTA's second tile (its latency hides under the first tile's setup)
  if (warp == 0) {
    tmem_alloc(holder, 256);
  } else if (tma_thread) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_ld, 1);
    mbar_init(bar_1, 1);
    mbar_init(bar_dv, 1);
    mbar_init(bar_2, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar_ld, 4u * box_bytes);
    tma_load_4d(sQ, &tm_qkv, bar_ld, x.h * HD, x.c1, x.c2, x.b);
    tma_load_4d(sK, &tm_qkv, bar_ld, a.H * HD + x.h * HD, x.c1, x.c2, x.b);
    tma_load_4d(sV, &tm_qkv, bar_ld, 2 * a.H * HD + x.h * HD, x.c1, x.c2, x.b);
    tma_load_4d(sDO, &tm_do, bar_ld, x.h * HD, x.c1, x.c2, x.b);
  } else if (warp == 3) {
    // rows [L, LP) of every tile are contraction rows of some MMA (queries for dV / dK, keys for dQ) or feed masked columns: zero them.
    // Nothing writes them again (the TMA boxes cover rows < `rows`, the CLS row is row `rows` = L - 1): once per CTA is enough.
    const int pad = a.LP - a.L;
    for (int i = lane; i < pad * 32; i += 32) {
      const int m = i / (pad * 8), rem = i - m * pad * 8;
      st_shared_v4(base + (uint32_t)m * ATOM + swz(a.L + (rem >> 3), rem & 7), 0u, 0u, 0u, 0u);
    }
  }
  // what a tile needs besides its TMA boxes (all of it outside the rows the TMA loads write, so it runs while the next tile's boxes are
  // landing): the CLS rows of q, k, v, dO -> row `rows` of the four tiles; and per row: the key mask, the token, the log-sum-exp of the
  // row's FULL key set (the CLS query's covers the whole clip) and, for the CLS query, delta = dO . O (its keys span every tile, so it
  // cannot come from this tile)
  uint32_t km[2];
  int tok = -1;                   // token (within the sample) of this thread's row; -1: none (padding, or the CLS row: merged separately)
  float row_lse = 0.f, delta = 0.f;
  auto setup = [&](const TileId& y) {
    if (warp == 2 && a.mode != 0) {
      const int m = lane >> 3, c = lane & 7;
      const bf16* src = m < 3 ? qkv + (long long)y.b * a.N * ld + (long long)m * a.H * HD + y.h * HD + c * 8
                              : dout + (long long)y.b * a.N * ldo + y.h * HD + c * 8;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
      st_shared_v4(base + (uint32_t)m * ATOM + swz(a.rows, c), v.x, v.y, v.z, v.w);
    }
    const unsigned long long kmask = key_mask(a, y.g, r, half * 64);
    km[0] = (uint32_t)kmask;
    km[1] = (uint32_t)(kmask >> 32);
    tok = r < a.rows ? row_token(a, y.g, r) : -1;
    row_lse = 0.f;
    delta = 0.f;
    const int ltok = cls_row ? 0 : tok;
    if (ltok >= 0) row_lse = lse[((long long)y.b * a.H + y.h) * a.N + ltok];
    if (cls_row) {
      const uint4* po = reinterpret_cast<const uint4*>(out + (long long)y.b * a.N * ldo + y.h * HD);
      const uint4* pd = reinterpret_cast<const uint4*>(dout + (long long)y.b * a.N * ldo + y.h * HD);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 p = __ldg(po + c), q = __ldg(pd + c);
        const float2 x0 = unpack_bf16x2(p.x), x1 = unpack_bf16x2(p.y), x2 = unpack_bf16x2(p.z), x3 = unpack_bf16x2(p.w);
        const float2 y0 = unpack_bf16x2(q.x), y1 = unpack_bf16x2(q.y), y2 = unpack_bf16x2(q.z), y3 = unpack_bf16x2(q.w);
        delta += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
      }
    }
  };
  // L2 prefetch of everything setup(y) reads with plain loads, one tile ahead of it (lanes 1..31 of warp 1)
  auto prefetch_plain = [&](const TileId& y) {
    const float* l2 = lse + ((long long)y.b * a.H + y.h) * a.N;
    if (a.mode != 0) {
      if (lane <= 3) prefetch_l2(qkv + (long long)y.b * a.N * ld + (long long)(lane - 1) * a.H * HD + y.h * HD);
      else if (lane == 4) prefetch_l2(dout + (long long)y.b * a.N * ldo + y.h * HD);
      else if (lane == 5) prefetch_l2(out + (long long)y.b * a.N * ldo + y.h * HD);
      else if (lane == 6) prefetch_l2(l2);
    }
    if (lane >= 8) {
      if (a.mode == 0) prefetch_l2_range(l2, a.N * 4, lane - 8, 24);
      else if (a.mode == 1) prefetch_l2_range(l2 + 1 + y.g * a.n, a.n * 4, lane - 8, 24);
      else
        for (int tt = lane - 8; tt < a.T; tt += 24) prefetch_l2_range(l2 + 1 + tt * a.n + y.g * a.GP, a.GP * 4, 0, 1);
    }
  };
  // CLS token: thread j < 192 publishes element j of a tile's partial dq (CLS query) / dk / dv (CLS key) and later takes a ticket of
  // element j; the thread that draws the LAST ticket of element j (any tile of this (b, h)) sums the partials in tile order -> row 0 of
  // dqkv (+ its bias gradient).  Per-element tickets, drawn half a tile after the partial was published: no barrier and no thread ever
  // waits for a store + fence + atomic round trip.
  auto cls_publish = [&](int bh, int g) { cls_ws[((long long)bh * a.chunks + g) * 192 + tid] = s_cls[tid]; };
  auto cls_ticket = [&](int bh) {
    __threadfence();                               // the partial is visible before the ticket is
    return atomicAdd(tickets + (long long)bh * 192 + tid, 1);
  };
  auto cls_merge = [&](int bh_) {
    const long long bh = bh_;
    tickets[bh * 192 + tid] = 0;                   // ready for the next launch (stream-ordered)
    __threadfence();
    const float* w = cls_ws + bh * a.chunks * 192 + tid;       // other SMs wrote these: read through L2
    float acc = 0.f;
#pragma unroll 8
    for (int gg = 0; gg < a.chunks; ++gg) acc += __ldcg(w + gg * 192);
    const int m = tid >> 6, d = tid & 63;
    const long long hcol = (long long)m * a.H * HD + (bh % a.H) * HD + d;
    const bf16 o = opnd_from_float(acc);
    dqkv[(bh / a.H) * a.N * ld + hcol] = o;
    if (dbias != nullptr && m != 1) {              // the CLS token's share of the bias gradient: its dq (q third), its dO row (v third)
      const bf16 c = m == 0 ? o : dout[(bh / a.H) * a.N * ldo + (bh % a.H) * HD + d];
#ifdef TVTS_OPERAND_FP16
      atomicAdd(dbias + hcol, __half2float(c));
#else
      atomicAdd(dbias + hcol, __bfloat162float(c));
#endif
    }
  };
  setup(x);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;
  const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)half * 64u;
  const uint32_t tq = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)half * 32u;
  unsigned tn = s_next;          // the next tile of this CTA (>= total: none)
  int pub_bh = -1, pub_g = 0;    // the tile whose CLS partial sits in s_cls (left there by the previous iteration)

  for (uint32_t it = 0; t < total; ++it) {
    const uint32_t ph = it & 1u;
    const bool has_next = tn < total;
    TileId y = x;
    if (has_next) y = decode_tile(a, tn);
    unsigned drawn = total;
    if (tid == 64 && has_next) drawn = G + (unsigned)atomicAdd(sched, 1);     // dynamic tile scheduler, two tiles ahead
    PROF_STAMP(0);
    if (a.ahead && has_next && warp == 1 && lane >= 1) prefetch_plain(y);
    if (tma_thread) {
      if (a.ahead && has_next) {       // the next tile's boxes -> L2: the TMA loads issued behind bar_1 / bar_dv / bar_2 then hit L2
        tma_prefetch_4d(&tm_qkv, 2 * a.H * HD + y.h * HD, y.c1, y.c2, y.b);
        tma_prefetch_4d(&tm_do, y.h * HD, y.c1, y.c2, y.b);
        tma_prefetch_4d(&tm_qkv, y.h * HD, y.c1, y.c2, y.b);
        tma_prefetch_4d(&tm_qkv, a.H * HD + y.h * HD, y.c1, y.c2, y.b);
      }
      mbar_wait(bar_ld, ph);
      tc_fence_after();
      const uint32_t idesc = make_idesc(a.LP, false, false);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc, k > 0 ? 1u : 0u);            // S = Q K^T
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + 128, desc_kmajor(sDO, k), desc_kmajor(sV, k), idesc, k > 0 ? 1u : 0u);    // dP = dO V^T
      umma_commit(bar_1);
    }
    __syncwarp();
    const int pend_bh = tid < 192 ? pub_bh : -1;
    if (pend_bh >= 0) cls_publish(pend_bh, pub_g);          // the previous tile's CLS partial, under the S / dP MMAs
    mbar_wait(bar_1, ph);
    tc_fence_after();
    PROF_STAMP(1);
    if (tma_thread && has_next) {      // V was last read by the dP MMAs: the next tile's box may land
      mbar_arrive_expect_tx(bar_ld, 4u * box_bytes);
      tma_load_4d(sV, &tm_qkv, bar_ld, 2 * a.H * HD + y.h * HD, y.c1, y.c2, y.b);
    }

    const float lse2 = row_lse * LOG2E;
    uint32_t pk[2][16];
    uint32_t v[32], w[32];
    // ---- pass A: P = exp(S * scale - lse) -> shared memory (A operand of dV = P^T dO) and registers; delta = sum_j P_ij dP_ij
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (half * 64 + c * 32 < a.LP) {       // warp-uniform
        tmem_ld_row32(trow + c * 32, v);
        tmem_ld_32x32(trow + 128 + c * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float p0 = 0.f, p1 = 0.f;
          if ((km[c] >> (2 * j)) & 1u) { p0 = exp2f(fmaf(__uint_as_float(v[2 * j]), sl2, -lse2)); dsum = fmaf(p0, __uint_as_float(w[2 * j]), dsum); }
          if ((km[c] >> (2 * j + 1)) & 1u) { p1 = exp2f(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -lse2)); dsum = fmaf(p1, __uint_as_float(w[2 * j + 1]), dsum); }
          pk[c][j] = pack_bf16x2(p0, p1);
        }
        const uint32_t atom = sP + (uint32_t)half * ATOM;
#pragma unroll
        for (int q = 0; q < 4; ++q) st_shared_v4(atom + swz(r, c * 4 + q), pk[c][4 * q], pk[c][4 * q + 1], pk[c][4 * q + 2], pk[c][4 * q + 3]);
      }
    }
    xch[half * 128 + r] = dsum;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();           // P complete; every thread is done with the S columns (dV overwrites [0,64))
    PROF_STAMP(2);

    if (tma_thread) {
      tc_fence_after();
      const uint32_t idesc = make_idesc(HD, true, true);
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tmem, desc_mnmajor(sP, ks), desc_mnmajor(sDO, ks), idesc, ks > 0 ? 1u : 0u);   // dV = P^T dO
      umma_commit(bar_dv);
    }
    __syncwarp();
    const float dlt = cls_row ? delta : dsum + xch[(half ^ 1) * 128 + r];
    // the ticket of the partial published at the top of this iteration: its round trip runs under the dV MMAs and pass B
    int ticket = -1;
    if (pend_bh >= 0) ticket = cls_ticket(pend_bh);
    if (dbias != nullptr) {
      // bias gradient of the v third of the qkv Linear.  Every softmax row sums to 1, so sum_j dV[j, :] = sum_i (sum_j P[i, j]) dO[i, :] =
      // the column sums of dO over the tokens: read from the dO tile while the dV MMAs run (warp w owns the 16-byte chunk w: a lane
      // sums every 32nd row, a transposing butterfly over the low three lane bits leaves lane l with column l & 7, two more steps finish).
      // The CLS token's dO row is added by the thread that merges the CLS partials.  (The k third is identically zero -- a constant added
      // to every key shifts all scores of a query alike -- so nothing is accumulated for it.)
      float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int rr = lane; rr < a.rows; rr += 32) {
        const uint4 u = *reinterpret_cast<const uint4*>(gen + 3 * ATOM + swz(rr, warp));
        const float2 x0 = unpack_bf16x2(u.x), x1 = unpack_bf16x2(u.y), x2 = unpack_bf16x2(u.z), x3 = unpack_bf16x2(u.w);
        s8[0] += x0.x; s8[1] += x0.y; s8[2] += x1.x; s8[3] += x1.y; s8[4] += x2.x; s8[5] += x2.y; s8[6] += x3.x; s8[7] += x3.y;
      }
      float s4[4], s2[2];
#pragma unroll
      for (int k = 0; k < 4; ++k) s4[k] = ((lane & 4) ? s8[4 + k] : s8[k]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? s8[k] : s8[4 + k], 4);
#pragma unroll
      for (int k = 0; k < 2; ++k) s2[k] = ((lane & 2) ? s4[2 + k] : s4[k]) + __shfl_xor_sync(0xffffffffu, (lane & 2) ? s4[k] : s4[2 + k], 2);
      float sum = ((lane & 1) ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, (lane & 1) ? s2[0] : s2[1], 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      if (lane < 8) atomicAdd(dbias + 2LL * a.H * HD + x.h * HD + warp * 8 + lane, sum);
    }
    mbar_wait(bar_dv, ph);     // the MMAs have read P: its buffer may now take dS
    tc_fence_after();
    PROF_STAMP(3);

    // ---- pass B: dS = P o (dP - delta) * scale -> the same buffer
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (half * 64 + c * 32 < a.LP) {
        tmem_ld_row32(trow + 128 + c * 32, w);
        tmem_ld_wait();
        uint32_t ds[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // no mask here: P is exactly 0 where the key is masked, and dP is finite there (real keys, or the zeroed rows [L, LP) of V)
          const float2 p = unpack_bf16x2(pk[c][j]);
          const float d0 = p.x * a.scale * (__uint_as_float(w[2 * j]) - dlt);
          const float d1 = p.y * a.scale * (__uint_as_float(w[2 * j + 1]) - dlt);
          ds[j] = pack_bf16x2(d0, d1);
        }
        const uint32_t atom = sP + (uint32_t)half * ATOM;
#pragma unroll
        for (int q = 0; q < 4; ++q) st_shared_v4(atom + swz(r, c * 4 + q), ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();           // dS complete; every thread is done with the dP columns (dQ overwrites [128,192))
    PROF_STAMP(4);

    if (tma_thread) {
      if (has_next) tma_load_4d(sDO, &tm_do, bar_ld, y.h * HD, y.c1, y.c2, y.b);      // dO: last read by the dV MMAs and the bias sums above
      tc_fence_after();
      const uint32_t idesc_k = make_idesc(HD, true, true), idesc_q = make_idesc(HD, false, true);
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tmem + 64, desc_mnmajor(sP, ks), desc_mnmajor(sQ, ks), idesc_k, ks > 0 ? 1u : 0u);   // dK = dS^T Q
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tmem + 128, desc_kmajor(sP, ks), desc_mnmajor(sK, ks), idesc_q, ks > 0 ? 1u : 0u);  // dQ = dS K
      umma_commit(bar_2);
    }
    __syncwarp();

    // ---- epilogue: row r of dV (key r), then -- once the last MMAs are done -- dQ (query r) and dK (key r): 16-bit, this thread's 32 of
    // the 64 columns straight to global memory.  dV goes first, WHILE the dK / dQ MMAs run (its accumulator has been complete since bar_dv).
    // The q third of the qkv Linear's bias gradient (column sums of dQ over the tokens; replaces a separate pass over dqkv) comes from
    // the same registers: a transposing butterfly over the warp's 32 rows leaves lane l with column l, one atomicAdd per (warp, column);
    // the CLS row is added by the thread that merges it.
    bf16* drow = dqkv + ((long long)x.b * a.N + (tok >= 0 ? tok : 0)) * ld + x.h * HD + half * 32;
    float* dbcol = dbias != nullptr ? dbias + x.h * HD + half * 32 + lane : nullptr;
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) {
      const int m = mi == 0 ? 2 : mi - 1;                              // dV, dQ, dK
      if (mi == 1) {
        PROF_STAMP(5);
        mbar_wait(bar_2, ph);                                          // every MMA has completed: Q, K and the dS atoms are free
        tc_fence_after();
        PROF_STAMP(6);
        if (tma_thread && has_next) {
          tma_load_4d(sQ, &tm_qkv, bar_ld, y.h * HD, y.c1, y.c2, y.b);
          tma_load_4d(sK, &tm_qkv, bar_ld, a.H * HD + y.h * HD, y.c1, y.c2, y.b);
        }
      }
      const uint32_t tcol = m == 0 ? 128u : (m == 1 ? 64u : 0u);
      tmem_ld_row32(tq + tcol, v);
      tmem_ld_wait();
      if (tok >= 0) {
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        st_global_v8(drow + (long long)m * a.H * HD, o);
        st_global_v8(drow + (long long)m * a.H * HD + 16, o + 8);
      }
      if (cls_row) {              // CLS partials (fp32): handed to 192 threads through shared memory (published in the next iteration)
#pragma unroll
        for (int j = 0; j < 32; ++j) s_cls[m * 64 + half * 32 + j] = __uint_as_float(v[j]);
      }
      if (dbias != nullptr && m == 0) {       // q third: column sums of dQ (v: from dO above; k: identically zero)
        float f16[16], f8[8], f4[4], f2[2];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float lo = tok >= 0 ? __uint_as_float(v[k]) : 0.f, hi = tok >= 0 ? __uint_as_float(v[16 + k]) : 0.f;
          f16[k] = ((lane & 16) ? hi : lo) + __shfl_xor_sync(0xffffffffu, (lane & 16) ? lo : hi, 16);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) f8[k] = ((lane & 8) ? f16[8 + k] : f16[k]) + __shfl_xor_sync(0xffffffffu, (lane & 8) ? f16[k] : f16[8 + k], 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) f4[k] = ((lane & 4) ? f8[4 + k] : f8[k]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? f8[k] : f8[4 + k], 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) f2[k] = ((lane & 2) ? f4[2 + k] : f4[k]) + __shfl_xor_sync(0xffffffffu, (lane & 2) ? f4[k] : f4[2 + k], 2);
        const float sum = ((lane & 1) ? f2[1] : f2[0]) + __shfl_xor_sync(0xffffffffu, (lane & 1) ? f2[0] : f2[1], 1);
        atomicAdd(dbcol + (long long)m * a.H * HD, sum);
      }
    }
    PROF_STAMP(7);
    if (has_next) setup(y);    // the next tile's CLS rows, masks, tokens and row statistics
    if (pend_bh >= 0 && ticket == a.chunks - 1) cls_merge(pend_bh);      // (one tile in `chunks`) this thread merges that (b, h)
    pub_bh = a.mode != 0 ? x.b * a.H + x.h : -1;
    pub_g = x.g;
    x = y;
    if (tid == 64) s_next = drawn;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    PROF_STAMP(8);
    PROF_END();
    t = tn;
    tn = s_next;
  }
  if (tid == 0 && atomicAdd(sched + 1, 1) == (int)G - 1) {      // the last CTA to leave re-arms the scheduler for the next launch
    sched[0] = 0;
    sched[1] = 0;
  }
  if (pub_bh >= 0 && tid < 192) {                               // the last tile's CLS partial
    cls_publish(pub_bh, pub_g);
    if (cls_ticket(pub_bh) == a.chunks - 1) cls_merge(pub_bh);
  }
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}
// 4-D view {column, position, frame, sample} of a [B*N, cols] 16-bit row-major token matrix, 128B swizzle, box = [64 columns, b1, b2, 1]:
//   mode 0:      {cols, N, 1, B}   -- the whole sequence of a sample
//   modes 1 / 2: {cols, n, T, B} over the PATCH tokens (base pointer skips the CLS row of sample 0; sample stride N rows)
int make_map(CUtensorMap* m, const void* ptr, const TcShape& a, long long cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return tvts_set_error(TVTS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  const cuuint64_t row = (cuuint64_t)cols * 2;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4];
  const char* base = reinterpret_cast<const char*>(ptr);
  if (a.mode == 0) {
    dims[0] = (cuuint64_t)cols; dims[1] = (cuuint64_t)a.N; dims[2] = 1; dims[3] = (cuuint64_t)a.B;
    strides[0] = row; strides[1] = row * a.N; strides[2] = row * a.N;
    box[0] = 64; box[1] = (cuuint32_t)a.N; box[2] = 1; box[3] = 1;
  } else {
    base += row;
    dims[0] = (cuuint64_t)cols; dims[1] = (cuuint64_t)a.n; dims[2] = (cuuint64_t)a.T; dims[3] = (cuuint64_t)a.B;
    strides[0] = row; strides[1] = row * a.n; strides[2] = row * a.N;
    box[0] = 64; box[1] = (cuuint32_t)(a.mode == 1 ? a.n : a.GP); box[2] = (cuuint32_t)(a.mode == 1 ? 1 : a.T); box[3] = 1;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType t16 = TVTS_OPERAND_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(m, t16, 4, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tvts_set_error(TVTS_ERR_CUDA, "attn_tc: cuTensorMapEncodeTiled failed (%d) mode=%d N=%d n=%d T=%d GP=%d", (int)r, a.mode, a.N, a.n, a.T, a.GP);
  return TVTS_OK;
}

bool make_shape(TcShape* s, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale) {
  if (d != HD || B <= 0 || H <= 0 || B * H * 192 > (int64_t)(kTicketBytes / sizeof(int)) - 64) return false;
  TcShape a{};
  a.B = (int)B; a.N = (int)N; a.H = (int)H; a.mode = (int)mode; a.T = (int)T; a.n = (int)n; a.causal = (int)causal; a.scale = scale;
  if (mode == 0) {
    if (N < 1 || N > 128) return false;
    a.T = 1; a.n = 0; a.GP = 0; a.chunks = 1; a.rows = (int)N; a.L = (int)N;
  } else if (mode == 1) {
    if (causal || T < 1 || n < 1 || n > 127 || N != 1 + T * n) return false;
    a.GP = 0; a.chunks = (int)T; a.rows = (int)n; a.L = (int)n + 1;
  } else if (mode == 2) {
    if (causal || T < 1 || T > 127 || n < 1 || N != 1 + T * n) return false;
    const int gp_max = 127 / (int)T;                     // T * GP patch rows + the CLS row fit one 128-row tile
    a.chunks = ((int)n + gp_max - 1) / gp_max;
    a.GP = ((int)n + a.chunks - 1) / a.chunks;           // balanced tiles (98 positions, T = 8: 7 tiles of 14)
    a.rows = (int)T * a.GP; a.L = a.rows + 1;
  } else {
    return false;
  }
  a.LP = (a.L + 15) / 16 * 16;
  *s = a;
  return true;
}

// per-device workspace: [8 MB of int tickets, 64 (forward) / 192 (backward) per (b, h), zero between launches] [CLS partials] (grown on demand; allocation happens
// outside any stream capture: the first call of a shape runs in the pre-capture warm-up step)
struct Workspace { float* ptr = nullptr; size_t bytes = 0; };
float* workspace(size_t bytes) {
  bytes += kTicketBytes;
  static Workspace per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Workspace& w = per_dev[dev];
  if (w.bytes < bytes) {
    // a buffer that is outgrown is NOT freed: CUDA graphs captured earlier keep launching kernels that hold its address
    w.ptr = nullptr; w.bytes = 0;
    size_t want = bytes < (32u << 20) ? (32u << 20) : bytes;
    if (cudaMalloc(&w.ptr, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(w.ptr, 0, kTicketBytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    w.bytes = want;
  }
  return w.ptr;
}

// tile-scheduler state {next tile, CTAs done} of a launch: the last 64 ints of the ticket area, one pair per kernel kind.  Launches that
// may overlap in time use different pairs: the text tower (mode 0) runs on a second stream next to the video tower (modes 1 / 2); two
// launches of the SAME kind must be stream-ordered (as they already had to be for the CLS tickets).
constexpr size_t kSchedInts = 64;
int* sched_slot(float* ws, bool bwd, int mode) {
  return reinterpret_cast<int*>(ws) + kTicketBytes / sizeof(int) - kSchedInts + 2 * ((bwd ? 2 : 0) + (mode == 0 ? 1 : 0));
}

int g_attn_tc_prefetch = 1;
int g_fwd_ahead_per_sm = 1;   // forward L2 prefetch distance in CTAs per SM (environment TVTS_ATTN_FWD_AHEAD).  Measured on the c3 shape, cold L2
                              // (GPU call 23): 1 -> 51.4 us, 2 -> 52.3 us, 4 (= one full wave of resident CTAs) -> 54.2 us   // L2 prefetch of the next wave's tiles (tvts_attn_set_tc bit 2 clears it: A/B measurements)
int g_attn_tc = -1;     // -1: not decided yet (environment TVTS_ATTN_TC=0 switches the tcgen05 path off for the whole process)
int g_attn_tc_time = -1;   // mode 2 separately (TVTS_ATTN_TC_TIME=0 keeps the warp-per-slot time kernels)
int attn_tc_on(int mode) {
  if (g_attn_tc < 0) {
    const char* e = getenv("TVTS_ATTN_TC");
    g_attn_tc = (e != nullptr && e[0] == '0') ? 0 : 1;
    const char* p = getenv("TVTS_ATTN_TC_PREFETCH");
    if (p != nullptr && p[0] == '0') g_attn_tc_prefetch = 0;
    const char* f = getenv("TVTS_ATTN_FWD_AHEAD");
    if (f != nullptr && atoi(f) > 0) g_fwd_ahead_per_sm = atoi(f);
  }
  if (g_attn_tc_time < 0) {
    const char* e = getenv("TVTS_ATTN_TC_TIME");
    g_attn_tc_time = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_attn_tc && (mode != 2 || g_attn_tc_time);
}

}  // namespace

// on: bit 0 = modes 0 / 1, bit 1 = mode 2 (time), bit 2 = NO L2 prefetch; tvts_attn_set_tc(3) = everything (default), 0 = mma.sync kernels
extern "C" int tvts_attn_set_tc(int on) {
  g_attn_tc = (on & 1) ? 1 : 0;
  g_attn_tc_time = (on & 2) ? 1 : 0;
  g_attn_tc_prefetch = (on & 4) ? 0 : 1;
  return TVTS_OK;
}

#ifdef TVTS_ATTN_PROF
// copies the stamps of the last launch: dst = long long [8192][13] (12 stamps + the SM id)
extern "C" int tvts_attn_tc_prof_read(void* dst) {
  return cudaMemcpyFromSymbol(dst, g_prof, sizeof(g_prof)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int tvts_attn_tc_supported(int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal) {
  TcShape s;
  return attn_tc_on((int)mode) && make_shape(&s, B, N, H, d, mode, T, n, causal, 1.0f) ? 1 : 0;
}

extern "C" int tvts_attn_tc_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T,
                                int64_t n, int64_t causal, float scale, void* stream) {
  TcShape a;
  TVTS_REQUIRE(make_shape(&a, B, N, H, d, mode, T, n, causal, scale), "attn_tc_fwd: unsupported shape (d=%lld mode=%lld N=%lld n=%lld)",
               (long long)d, (long long)mode, (long long)N, (long long)n);
  TVTS_REQUIRE(qkv && out && lse, "attn_tc_fwd: null pointer");
  TVTS_REQUIRE(B * N < (1ll << 31) && B * H * a.chunks < (1ll << 31) && B * H * 192 <= (int64_t)(kTicketBytes / sizeof(int)) - 64, "attn_tc_fwd: too many rows / (b, h) pairs");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tq, to;
  int rc = make_map(&tq, qkv, a, 3 * H * HD);
  if (rc) return rc;
  rc = make_map(&to, out, a, H * HD);
  if (rc) return rc;
  float* ws = workspace((size_t)B * H * a.chunks * 192 * sizeof(float));
  TVTS_REQUIRE(ws != nullptr, "attn_tc_fwd: workspace allocation failed");
  static bool attr = false;
  if (!attr) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  // L2 prefetch distance in tiles: a quarter of the CTAs resident on the chip (four per SM: 51 KB of shared memory, 128 TMEM columns,
  // <= 64 registers each) measured best, see g_fwd_ahead_per_sm
  a.ahead_tiles = g_attn_tc_prefetch ? g_fwd_ahead_per_sm * tvts_num_sms() : 0;
  a.ahead = 0;
  dim3 grid((unsigned)a.chunks, (unsigned)H, (unsigned)B);
  attn_tc_fwd_kernel<<<grid, kTcThreads, FWD_SMEM, st>>>(tq, to, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), lse,
                                                         a.mode != 0 ? ws + kTicketBytes / sizeof(float) : nullptr, reinterpret_cast<int*>(ws), a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_attn_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int64_t B, int64_t N,
                                int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale, void* stream) {
  return tvts_attn_tc_bwd_bias(qkv, out, dout, lse, dqkv, nullptr, B, N, H, d, mode, T, n, causal, scale, stream);
}

extern "C" int tvts_attn_tc_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias, int64_t B,
                                     int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale,
                                     void* stream) {
  TcShape a;
  TVTS_REQUIRE(make_shape(&a, B, N, H, d, mode, T, n, causal, scale), "attn_tc_bwd: unsupported shape (d=%lld mode=%lld N=%lld n=%lld)",
               (long long)d, (long long)mode, (long long)N, (long long)n);
  TVTS_REQUIRE(qkv && out && dout && lse && dqkv, "attn_tc_bwd: null pointer");
  TVTS_REQUIRE(B * N < (1ll << 31) && B * H * a.chunks < (1ll << 31) && B * H * 192 <= (int64_t)(kTicketBytes / sizeof(int)) - 64, "attn_tc_bwd: too many rows / (b, h) pairs");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TVTS_REQUIRE((uintptr_t)dqkv % 32 == 0, "attn_tc_bwd: dqkv must be 32-byte aligned (32-byte register stores)");
  CUtensorMap tq, tdo;
  int rc = make_map(&tq, qkv, a, 3 * H * HD);
  if (rc) return rc;
  rc = make_map(&tdo, dout, a, H * HD);
  if (rc) return rc;
  float* ws = workspace((size_t)B * H * a.chunks * 192 * sizeof(float));
  TVTS_REQUIRE(ws != nullptr, "attn_tc_bwd: workspace allocation failed");
  static bool attr = false;
  if (!attr) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  const long long resident = 2LL * tvts_num_sms();     // two CTAs per SM by construction (98 KB of shared memory, 256 TMEM columns each)
  a.ahead = g_attn_tc_prefetch;
  a.ahead_tiles = 0;
  const long long total = (long long)B * H * a.chunks;
  dim3 grid((unsigned)(total < resident ? total : resident));      // persistent: one CTA per resident slot
  attn_tc_bwd_kernel<<<grid, kTcThreads, BWD_SMEM, st>>>(tq, tdo, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(out),
                                                         reinterpret_cast<const bf16*>(dout), lse,
                                                         ws + kTicketBytes / sizeof(float), dbias, reinterpret_cast<bf16*>(dqkv),
                                                         reinterpret_cast<int*>(ws), sched_slot(ws, true, a.mode), a);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
