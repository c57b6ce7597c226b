// tcgen05 GEMM for the TVTS hot path: every nn.Linear / projection of the video tower, text tower and
// sort head, in forward, dgrad and wgrad form (reference call sites
// v2/model/video_encoder_ViT_B_16.py:41,74,105-109,233).
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] )       bf16 operands, fp32 accumulation in TMEM
//
// Operand layouts (either operand independently):
//   K-major : X[rows, K] row-major (the contraction index is contiguous)    -> forward / dgrad
//   MN-major: X[K, rows] row-major (the row index is contiguous)            -> wgrad straight from
//             the [tokens, features] activations with tokens as the contraction, no transposes
//
// Structure: persistent CTAs (one per SM), warp-specialised:
//   warp 0  TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1  MMA issuer    (one thread, tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction)
//   work units = output tiles x K splits, ordered tile-fastest: units running concurrently cover all output tiles of one K range and
//           share its operand slices through L2 (wgrad: small output, K = 25k tokens)
//   warp 2  TMEM allocator (2 x BN fp32 columns: double-buffered accumulator so the epilogue of tile i
//           overlaps the MMAs of tile i+1)
//   warps 4-11 epilogue   (tcgen05.ld 32x32b: one accumulator row per thread; fused bias / QuickGELU|GELU / activation-derivative /
//           residual add in registers; the warp's [32 rows x 128 B] block goes through a 128B-swizzled shared-memory box to the TMA
//           engine: cp.async.bulk.tensor store, or cp.reduce.async.bulk .add for split-K / gradient accumulation)
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;

// CTAS = 1: one CTA per 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, cta_group::2) per 256 x BN tile -- each CTA stages its
// own 128 rows of A and HALF of the B tile (BN/2 rows), so operand shared-memory traffic per SM drops from 96 to 64 B/clk and the
// smaller stage buys a deeper TMA pipeline.
template <int BN, int CTAS>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = BN / CTAS;
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = kEpiWarps * 2 * 4096;          // per epilogue warp: ring of two [32 rows x 128 B] TMA-store boxes
  static constexpr int BIAS_BYTES = BN * 4;                       // the tile's bias slice
  static constexpr int STAGES = (227 * 1024 - EPI_BYTES - BIAS_BYTES - 512 - 1024) / STAGE_BYTES > 6
                                    ? 6 : (227 * 1024 - EPI_BYTES - BIAS_BYTES - 512 - 1024) / STAGE_BYTES;   // pair: 5, solo: 3 (BN=256) / 5 (BN=128)
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int BAR_OFFSET = BIAS_OFFSET + BIAS_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;  // barriers + tmem ptr + alignment slack
};

struct EpiParams {
  void* out;
  void* out_pre;
  const float* bias;
  const float* residual;
  const bf16* aux;
  long long ldo, ldr, ldaux;
  int out_dtype;   // 0 fp32, 1 bf16
  int act;         // applied to acc + bias
  int dact;        // multiply by act'(aux) (dgrad through an activation)
  int accumulate;  // 0 store, 1 red.add into fp32 out (split-K / grad accumulation)
  int opnd_mode;   // 0: residual / aux read straight from global memory; 1: residual, 2: aux prefetched by TMA into the store ring
  int direct;      // 1: outputs leave straight from registers (32-byte stores per thread), no shared-memory box / TMA store (see the epilogue)
  float alpha;
};

// 32 bytes (8 registers) to global memory in one instruction (STG.256, sm_100+); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void st_global_v4(void* p, const uint32_t* w) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
// this thread's `nwords` 32-bit words (a multiple of 8) of one output row, of which the first `valid` (a multiple of 4) lie inside the matrix
__device__ __forceinline__ void store_row_words(void* p, const uint32_t* w, int nwords, int valid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (8 * j < nwords) {
      if (valid >= 8 * j + 8) st_global_v8(reinterpret_cast<uint32_t*>(p) + 8 * j, w + 8 * j);
      else if (valid >= 8 * j + 4) st_global_v4(reinterpret_cast<uint32_t*>(p) + 8 * j, w + 8 * j);
    }
  }
}

struct GemmShape {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split, tiles;
};

template <int BN, bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_pre,
            const __grid_constant__ CUtensorMap tmap_opnd, GemmShape s, EpiParams ep, int dbg_lbo, int dbg_sbo, int dbg_kadv, int dbg_epi) {
  using L = SmemLayout<BN, CTAS>;
  constexpr int kStages = L::STAGES;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;   // rank inside the pair; rank 0 = leader (issues the MMAs)
  const int worker = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (kStages + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 2 + i); };
  const uint32_t tmem_holder = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_holder_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (2 * kStages + 4));
  auto opnd_bar = [&](int ew, int b) { return bar_base + 8u * (2 * kStages + 6 + 2 * ew + b); };   // epilogue-operand boxes (2 per warp)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (ep.opnd_mode != 0) tma_prefetch_desc(&tmap_opnd);
    if (ep.out_pre != nullptr) tma_prefetch_desc(&tmap_pre);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), CTAS * kEpiWarps);   // one elected lane per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(opnd_bar(i >> 1, i & 1), 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CTAS == 2) tmem_alloc_pair(tmem_holder, 2 * BN);
    else tmem_alloc(tmem_holder, 2 * BN);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything is signalled remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder_gen;
  // Programmatic dependent launch: the launch below carries cudaLaunchAttributeProgrammaticStreamSerialization, so this grid may be
  // scheduled while the previous kernel of the stream is still draining -- everything above (tensor-map prefetch, barrier init, TMEM
  // allocation, cluster sync) overlaps that tail; nothing above touches memory the previous kernel writes.  From here on it does.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // ... and let a dependent GEMM grid be scheduled as soon as this grid's CTAs start to retire (persistent grid: every CTA is resident
  // already, so the dependents only ever take SMs this grid has left; they block at their own griddepcontrol.wait until it completes)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int total_units = s.m_tiles * s.n_tiles * s.splits;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = worker; unit < total_units; unit += n_workers) {
      const int split = unit / s.tiles;
      const int tile = unit % s.tiles;
      const int n_blk = tile % s.n_tiles;
      const int m_blk = tile / s.n_tiles;
      const int kb0 = split * s.kb_per_split;
      const int kb1 = min(kb0 + s.kb_per_split, s.k_blocks);
      const int m0 = m_blk * (BLOCK_M * CTAS) + (int)cta_rank * BLOCK_M;   // this CTA's rows of A
      const int n0 = n_blk * BN + (int)cta_rank * L::B_ROWS;               // this CTA's rows of B
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        // pair: both CTAs' bytes complete on the LEADER's barrier (its MMA thread consumes both halves)
        const uint32_t fb = CTAS == 2 ? mapa_shared(full_bar(stage), 0) : full_bar(stage);
        if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), CTAS * L::STAGE_BYTES);
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
        auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1) {
          if (CTAS == 2) tma_load_2d_pair(dst, tm, fb, c0, c1);
          else tma_load_2d(dst, tm, fb, c0, c1);
        };
        if (!A_MN) {
          load(sa, &tmap_a, kb * BLOCK_K, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j) load(sa + j * (64 * BLOCK_K * 2), &tmap_a, m0 + j * 64, kb * BLOCK_K);
        }
        if (!B_MN) {
          load(sb, &tmap_b, kb * BLOCK_K, n0);
        } else {
#pragma unroll
          for (int j = 0; j < L::B_ROWS / 64; ++j) load(sb + j * (64 * BLOCK_K * 2), &tmap_b, n0 + j * 64, kb * BLOCK_K);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0 && cta_rank == 0) {
    // ===================== MMA issuer (single thread; the leader CTA of a pair) =====================
    // instruction descriptor: D=f32 (bit4), A / B format at bits 7 / 10 (kind::f16: 0 = f16, 1 = bf16), major bits 15/16, N>>3 at 17, M>>4 at 24
    constexpr uint32_t kOperandFormat = TVTS_OPERAND_IS_FP16 ? 0u : ((1u << 7) | (1u << 10));
    const uint32_t idesc = (1u << 4) | kOperandFormat | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BLOCK_M * CTAS) >> 4) << 24);
    // K-major SW128: 8-row groups are 1024 B apart (SBO); LBO unused.  Advance 32 B per UMMA_K inside the row.
    // MN-major SW128: 64-element MN groups are one TMA box (8192 B) apart (LBO); 8-k-row groups 1024 B apart (SBO);
    //                 UMMA_K = 16 k-rows = 2 groups -> advance 2048 B.
    const uint32_t a_lbo = A_MN ? (dbg_lbo ? dbg_lbo : 64 * BLOCK_K * 2) : 16;
    const uint32_t a_sbo = A_MN ? (dbg_sbo ? dbg_sbo : 1024) : 1024;
    const uint32_t a_adv = A_MN ? (dbg_kadv ? dbg_kadv : 2048) : UMMA_K * 2;
    const uint32_t b_lbo = B_MN ? (dbg_lbo ? dbg_lbo : 64 * BLOCK_K * 2) : 16;
    const uint32_t b_sbo = B_MN ? (dbg_sbo ? dbg_sbo : 1024) : 1024;
    const uint32_t b_adv = B_MN ? (dbg_kadv ? dbg_kadv : 2048) : UMMA_K * 2;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int unit = worker; unit < total_units; unit += n_workers, ++it) {
      const int split = unit / s.tiles;
      const int kb0 = split * s.kb_per_split;
      const int kb1 = min(kb0 + s.kb_per_split, s.k_blocks);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint64_t da = umma_smem_desc(sa + k * a_adv, a_lbo, a_sbo);
          const uint64_t db = umma_smem_desc(sb + k * b_adv, b_lbo, b_sbo);
          if (CTAS == 2) umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          else umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        if (CTAS == 2) umma_commit_pair(empty_bar(stage));   // frees the stage in BOTH CTAs
        else umma_commit(empty_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (CTAS == 2) umma_commit_pair(tfull_bar(acc));
      else umma_commit(tfull_bar(acc));
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps =====================
    // TMEM lane quadrant q = warp % 4 (hardware restriction); the two warps of a quadrant split the tile's columns.
    // tcgen05.ld gives every thread one accumulator ROW; the fused math (alpha, bias, activation, act', residual) runs in that
    // layout, then the warp writes its [32 rows x 128 B] block (64 bf16 or 32 fp32 columns) into a 128B-swizzled shared-memory
    // box (conflict-free 16-byte stores) and ONE lane hands it to the TMA engine: cp.async.bulk.tensor store (or
    // cp.reduce.async.bulk .add for split-K / gradient accumulation).  Stores are asynchronous, fully coalesced, clipped at the
    // matrix edges by the tensor map, and double-buffered per warp.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int ew = warp - 4;
    constexpr int COLS_PER_WARP = BN / 2;
    const uint32_t ring = smem_base + L::EPI_OFFSET + (uint32_t)ew * 8192u;
    const float* bias_s = reinterpret_cast<const float*>(smem_gen + L::BIAS_OFFSET);
    const uint32_t tempty_leader0 = CTAS == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    const bool out_bf16 = ep.out_dtype == 1;
    const int CH = out_bf16 ? 64 : 32;                 // columns per 128-byte box row
    uint32_t nstore = 0;                               // boxes issued by this warp (ring index; lane 0 owns the bulk groups)
    // wait until this warp's ring buffer(s) have been read out by the TMA engine (lane 0 issued them)
    auto box_acquire = [&](bool both) {
      if (lane == 0) { if (both) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
      __syncwarp();
    };
    // write `nchunks` 16-byte chunks (starting at chunk0) of this thread's 128-byte row into a 128B-swizzled box
    auto box_write = [&](uint32_t buf, int chunk0, const uint32_t* w, int nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nchunks) {
          const uint32_t addr = buf + (uint32_t)lane * 128u + (uint32_t)(((chunk0 + j) ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]),
                       "r"(w[4 * j + 3])
                       : "memory");
        }
      }
    };
    auto box_issue = [&](const CUtensorMap* tm, uint32_t buf, int col0, int row0, bool reduce) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (reduce) tma_reduce_add_2d(tm, buf, col0, row0);
        else tma_store_2d(tm, buf, col0, row0);
        bulk_commit();
      }
    };
    const bool has_pre = ep.out_pre != nullptr;
    // Direct mode (opt-in, tvts_gemm_set_epilogue_direct): the results leave straight from registers -- every thread owns one output row,
    // 32 consecutive columns per pass = 64 (bf16) or 128 (fp32) contiguous bytes, written with 32-byte stores (whole sectors).  No
    // shared-memory box, no proxy fence, no TMA store whose read-out gates the reuse of a two-box ring; the epilogue operand (residual /
    // act' input) is still prefetched by TMA into the ring, whose boxes are then free again as soon as they are read.  Built to test
    // whether the store path bounds the K = 768 problems (qkv, c_fc with two outputs, proj + residual): it does not (see g_epi_direct).
    const bool direct = ep.direct != 0;
    // Epilogue operand (the fp32 residual, or the bf16 pre-activation whose act' multiplies a dgrad) is PREFETCHED by TMA into
    // the very ring box the result will be stored from ([32 rows x 128 B] = 32 fp32 / 64 bf16 columns, same shape as the output
    // box): the first two boxes of a tile are requested before the accumulator is even ready, the following ones as soon as the
    // box two chunks back has been read out.  The thread reads its row from the swizzled box, combines, and overwrites it in place.
    const bool has_opnd = ep.opnd_mode != 0;
    uint32_t nload = 0;                                 // operand boxes requested / consumed by this warp (ring + phase index)
    uint32_t nuse = 0;
    auto opnd_request = [&](int col0, int row0) {       // lane 0 only
      const uint32_t b = nload & 1u;
      mbar_arrive_expect_tx(opnd_bar(ew, b), 4096u);
      tma_load_2d(ring + b * 4096u, &tmap_opnd, opnd_bar(ew, b), col0, row0);
    };
    const int nchunks = COLS_PER_WARP / CH;
    int it = 0;
    for (int unit = worker; unit < total_units; unit += n_workers, ++it) {
      const int tile = unit % s.tiles;
      const bool first_split = (unit / s.tiles) == 0;  // bias / residual are added by one split only
      const int n_blk = tile % s.n_tiles;
      const int m_blk = tile / s.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool use_res = ep.residual != nullptr && first_split;
      const bool use_bias = ep.bias != nullptr && first_split;
      const int row0 = m_blk * (BLOCK_M * CTAS) + (int)cta_rank * BLOCK_M + q * 32;
      const long long grow = (long long)row0 + lane;
      const bool row_ok = grow < s.M;
      const int col_base = n_blk * BN + half * COLS_PER_WARP;
      const bool tile_opnd = has_opnd && (ep.opnd_mode == 2 || use_res);
      // operand boxes of the first two chunks: requested while the MMAs of this tile are still running
      if (tile_opnd) {
        if (lane == 0) {
          bulk_wait_read<0>();                           // both ring boxes have been read out by earlier stores
          for (int c = 0; c < 2 && c < nchunks; ++c)
            if (col_base + c * CH < s.N) { opnd_request(col_base + c * CH, row0); ++nload; }
        }
        nload = __shfl_sync(0xffffffffu, nload, 0);
      }
      // stage the tile's bias slice (all 8 epilogue warps = 256 threads, one column each); named barrier 1 = epilogue warps only
      if (ep.bias != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the previous tile's readers are done
        const int t = (int)threadIdx.x - 128;
        if (use_bias && t < BN) {
          const int col = n_blk * BN + t;
          reinterpret_cast<float*>(smem_gen + L::BIAS_OFFSET)[t] = col < s.N ? __ldg(ep.bias + col) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * COLS_PER_WARP;
      int pend_col = -1;
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += CH) {
        const int col0 = col_base + c0;
        if (col0 >= s.N || dbg_epi == 3) break;  // warp-uniform
        uint32_t buf_o = ring, buf_p = 0;
        if (tile_opnd) {
          buf_o = ring + (nuse & 1u) * 4096u;            // the operand box of this chunk; the result overwrites it (TMA-store mode)
          mbar_wait(opnd_bar(ew, nuse & 1u), (nuse >> 1) & 1u);
          ++nuse;
        } else if (direct) { }                           // no output box
        else if (has_pre) { buf_p = ring; buf_o = ring + 4096u; }
        else { buf_o = ring + (nstore & 1u) * 4096u; ++nstore; }
        // the ring box(es) of this chunk are acquired (= their previous TMA store has been read out) only right before the first write
        // below, i.e. AFTER the chunk's first TMEM load and bias add: with two outputs per chunk (c_fc: pre-activation + activation)
        // both boxes are still being read out when the chunk starts, and that wait used to be fully exposed
        bool acquired = tile_opnd || direct;
#pragma unroll 1
        for (int sub = 0; sub < CH; sub += 32) {   // 32 accumulator columns per pass
          const int cs = col0 + sub;
          uint32_t v[32];
          tmem_ld_32x32(t_base + c0 + sub, v);
          // operand of this pass: from the prefetched box (swizzled, conflict-free), else straight from global memory
          float4 r4[8];
          uint4 x4[4];
          if (tile_opnd) {
            if (ep.opnd_mode == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t addr = buf_o + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r4[j].x), "=f"(r4[j].y), "=f"(r4[j].z), "=f"(r4[j].w) : "r"(addr));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t addr = buf_o + (uint32_t)lane * 128u + (uint32_t)((((sub >> 3) + j) ^ (lane & 7)) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x4[j].x), "=r"(x4[j].y), "=r"(x4[j].z), "=r"(x4[j].w) : "r"(addr));
              }
            }
          } else {
            if (use_res && row_ok) {
              const float* rp = ep.residual + grow * ep.ldr + cs;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                r4[j] = (cs + 4 * j < s.N) ? *reinterpret_cast<const float4*>(rp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (ep.dact != TVTS_ACT_NONE && row_ok) {
              const bf16* ap = ep.aux + grow * ep.ldaux + cs;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                x4[j] = (cs + 8 * j < s.N) ? *reinterpret_cast<const uint4*>(ap + 8 * j) : make_uint4(0, 0, 0, 0);
            }
          }
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * ep.alpha;
          if (use_bias) {
            const float* bp = bias_s + half * COLS_PER_WARP + c0 + sub;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(bp + 4 * j);   // shared-memory broadcast
              f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
            }
          }
          if (!acquired) { box_acquire(has_pre); acquired = true; }      // warp-uniform
          if (pend_col >= 0) {                                           // warp-uniform: the operand box two chunks ahead (see below)
            if (lane == 0) { bulk_wait_read<0>(); opnd_request(pend_col, row0); }
            ++nload;
            pend_col = -1;
          }
          if (has_pre) {                               // bf16 copy of the pre-activation
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
            if (!direct) box_write(buf_p, sub >> 3, v, 4);
            else if (row_ok) store_row_words(reinterpret_cast<bf16*>(ep.out_pre) + grow * ep.ldo + cs, v, 16, min(s.N - cs, 32) >> 1);
          }
          if (ep.act != TVTS_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = act_fwd(f[j], ep.act);
          }
          if (ep.dact != TVTS_ACT_NONE && (row_ok || tile_opnd)) {     // (out-of-range rows of a prefetched box are zero-filled)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 p0 = unpack_bf16x2(x4[j].x), p1 = unpack_bf16x2(x4[j].y), p2 = unpack_bf16x2(x4[j].z), p3 = unpack_bf16x2(x4[j].w);
              f[8 * j] *= act_bwd(p0.x, ep.dact); f[8 * j + 1] *= act_bwd(p0.y, ep.dact);
              f[8 * j + 2] *= act_bwd(p1.x, ep.dact); f[8 * j + 3] *= act_bwd(p1.y, ep.dact);
              f[8 * j + 4] *= act_bwd(p2.x, ep.dact); f[8 * j + 5] *= act_bwd(p2.y, ep.dact);
              f[8 * j + 6] *= act_bwd(p3.x, ep.dact); f[8 * j + 7] *= act_bwd(p3.y, ep.dact);
            }
          }
          if (use_res && (row_ok || tile_opnd)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { f[4 * j] += r4[j].x; f[4 * j + 1] += r4[j].y; f[4 * j + 2] += r4[j].z; f[4 * j + 3] += r4[j].w; }
          }
          if (out_bf16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
            if (!direct) box_write(buf_o, sub >> 3, v, 4);          // 32 bf16 columns = 64 bytes = 4 chunks
            else if (row_ok) store_row_words(reinterpret_cast<bf16*>(ep.out) + grow * ep.ldo + cs, v, 16, min(s.N - cs, 32) >> 1);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(f[j]);
            if (!direct) box_write(buf_o, 0, v, 8);                 // 32 fp32 columns = the whole 128-byte row
            else if (row_ok) store_row_words(reinterpret_cast<float*>(ep.out) + grow * ep.ldo + cs, v, 32, min(s.N - cs, 32));
          }
        }
        if (dbg_epi != 1 && !direct) {
          if (has_pre) box_issue(&tmap_pre, buf_p, col0, row0, false);
          box_issue(&tmap_out, buf_o, col0, row0, ep.accumulate != 0);
        }
        if (tile_opnd) {
          // request the operand box two chunks ahead: it reuses THIS chunk's box ...
          const int cn = c0 + 2 * CH;
          const bool more = cn < COLS_PER_WARP && col_base + cn < s.N;     // warp-uniform
          if (more) {
            if (direct) {          // ... right away: every lane has read its row of the box into registers, nothing else touches it
              __syncwarp();
              if (lane == 0) opnd_request(col_base + cn, row0);
              ++nload;
            } else {
              pend_col = col_base + cn;      // ... from the NEXT chunk, after its TMEM load: by then the store just issued has been read out
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);   // the leader's MMA thread waits for both CTAs
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (lane == 0) bulk_wait_all();   // every box has left shared memory and been written before the CTA may exit
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or remote arrives are in flight
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D tensor map: inner (contiguous) extent d0, outer extent d1, outer stride ld elements of `esize` bytes; 128B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, long long d0, long long d1, long long ld, int box0, int box1, int esize = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return tvts_set_error(TVTS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType t16 = TVTS_OPERAND_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(m, esize == 2 ? t16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tvts_set_error(TVTS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p dims=%lld,%lld ld=%lld box=%d,%d esize=%d", (int)r, ptr,
                          d0, d1, ld, box0, box1, esize);
  return TVTS_OK;
}

int g_dbg_lbo = 0, g_dbg_sbo = 0, g_dbg_kadv = 0, g_dbg_epi = 0;
int g_pair_mode = -1;   // -1 auto, 0 never use CTA pairs, 1 always when the shape allows
int g_opnd_prefetch = 1; // TMA prefetch of the residual / aux epilogue operand
int g_solo_penalty = -1; // see tvts_gemm: solo-vs-pair tile choice against wave quantisation (-1: read the environment)
int g_epi_direct = -1;   // outputs stored straight from registers instead of through TMA-store boxes (environment TVTS_GEMM_EPI_DIRECT=1: on).
                         // OFF by default: measured neutral on the c3 step (round 2, profiles/r2/call24_*: GEMM time 22.15 vs 22.15 ms, c_fc with
                         // two outputs +3.5 %, proj + fp32 residual -4.6 %): those epilogues are bound by the bytes they move, not by the store path
int g_pdl = -1;          // programmatic dependent launch of the GEMM grids (environment TVTS_GEMM_PDL=0 turns it off)

template <int BN, bool A_MN, bool B_MN, int CTAS>
int launch(const tvts_gemm_args* g, const GemmShape& s, const EpiParams& ep, cudaStream_t stream) {
  using L = SmemLayout<BN, CTAS>;
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN) rc = make_tmap(&ta, g->a, g->K, g->M, g->lda, BLOCK_K, BLOCK_M);
  else rc = make_tmap(&ta, g->a, g->M, g->K, g->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&tb, g->b, g->K, g->N, g->ldb, BLOCK_K, L::B_ROWS);
  else rc = make_tmap(&tb, g->b, g->N, g->K, g->ldb, 64, BLOCK_K);
  if (rc) return rc;
  // epilogue boxes: [32 rows x 128 bytes] of the output (and of the optional bf16 pre-activation copy)
  CUtensorMap tout, tpre;
  const int oes = g->out_dtype == 1 ? 2 : 4;
  rc = make_tmap(&tout, g->out, g->N, g->M, g->ldo, 128 / oes, 32, oes);
  if (rc) return rc;
  if (g->out_pre) {
    rc = make_tmap(&tpre, g->out_pre, g->N, g->M, g->ldo, 64, 32, 2);
    if (rc) return rc;
  } else {
    tpre = tout;
  }
  // epilogue operand prefetch (see the kernel): residual with an fp32 output, or act'(aux) with a bf16 output
  CUtensorMap topnd = tout;
  EpiParams ep2 = ep;
  ep2.opnd_mode = 0;
  if (g_opnd_prefetch && !g->out_pre) {
    if (g->residual && !g->dact && g->out_dtype == 0) {
      rc = make_tmap(&topnd, g->residual, g->N, g->M, g->ldr, 32, 32, 4);
      if (rc) return rc;
      ep2.opnd_mode = 1;
    } else if (g->dact && !g->residual && g->out_dtype == 1) {
      rc = make_tmap(&topnd, g->aux, g->N, g->M, g->ldaux, 64, 32, 2);
      if (rc) return rc;
      ep2.opnd_mode = 2;
    }
  }
  ep2.direct = 0;
  if (g_epi_direct < 0) {
    const char* e = getenv("TVTS_GEMM_EPI_DIRECT");
    g_epi_direct = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (g_epi_direct && !g->accumulate && (uintptr_t)g->out % 32 == 0 && (g->ldo * oes) % 32 == 0 &&
      (!g->out_pre || (uintptr_t)g->out_pre % 32 == 0))
    ep2.direct = 1;
  auto kern = gemm_kernel<BN, A_MN, B_MN, CTAS>;
  if (g_pdl < 0) {
    const char* e = getenv("TVTS_GEMM_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    TVTS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int units = s.m_tiles * s.n_tiles * s.splits;
  const int workers = tvts_num_sms() / CTAS;
  const int grid = (units < workers ? units : workers) * CTAS;
  int prof_slot;
  tvts_prof_begin(stream, 2.0 * (double)s.M * (double)s.N * (double)s.K, 0.0, &prof_slot);
  tvts_prof_tag(prof_slot, s.M, s.N, s.K,
                (A_MN ? 1 : 0) | (B_MN ? 2 : 0) | (CTAS == 2 ? 4 : 0) | (ep.out_dtype ? 8 : 0) | (ep.residual ? 16 : 0) | (ep.out_pre ? 32 : 0) |
                    (ep.dact ? 64 : 0) | (ep.accumulate ? 128 : 0) | (s.splits << 8));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, tpre, topnd, s, ep2, g_dbg_lbo, g_dbg_sbo, g_dbg_kadv, g_dbg_epi);
  tvts_prof_end(stream, prof_slot);
  tvts_count_launch(1);
  if (le != cudaSuccess) return tvts_set_error(TVTS_ERR_CUDA, "gemm launch failed: %s", cudaGetErrorString(le));
  return TVTS_OK;
}

}  // namespace

// how many clusters of `cluster_size` CTAs of the pair kernel can be co-resident (GPC packing check for multicast designs)
extern "C" int tvts_gemm_debug_max_clusters(int cluster_size) {
  auto kern = gemm_kernel<256, false, false, 2>;
  using L = SmemLayout<256, 2>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 * 4);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}
extern "C" int tvts_gemm_set_pair_mode(int mode) {
  g_pair_mode = mode;
  return TVTS_OK;
}
extern "C" int tvts_gemm_set_epilogue_direct(int on) {
  g_epi_direct = on ? 1 : 0;
  return TVTS_OK;
}
extern "C" int tvts_gemm_set_operand_prefetch(int on) {
  g_opnd_prefetch = on;
  return TVTS_OK;
}
extern "C" int tvts_gemm_debug_epi(int mode) {
  g_dbg_epi = mode;
  return TVTS_OK;
}
extern "C" int tvts_gemm_debug_set(int lbo, int sbo, int kadv) {
  g_dbg_lbo = lbo; g_dbg_sbo = sbo; g_dbg_kadv = kadv;
  return TVTS_OK;
}

extern "C" int tvts_gemm(const tvts_gemm_args* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TVTS_REQUIRE(g != nullptr, "tvts_gemm: null args");
  TVTS_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "tvts_gemm: empty problem M=%lld N=%lld K=%lld", g->M, g->N, g->K);
  TVTS_REQUIRE(g->M < (1ll << 31) && g->N < (1ll << 31) && g->K < (1ll << 31), "tvts_gemm: dims exceed int32");
  TVTS_REQUIRE(g->a && g->b && g->out, "tvts_gemm: null operand pointer");
  TVTS_REQUIRE(g->N % 4 == 0, "tvts_gemm: N=%lld must be a multiple of 4", g->N);
  TVTS_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "tvts_gemm: operand leading dims must be multiples of 8 elements (lda=%lld ldb=%lld)",
               g->lda, g->ldb);
  TVTS_REQUIRE(((uintptr_t)g->a % 16 == 0) && ((uintptr_t)g->b % 16 == 0) && ((uintptr_t)g->out % 16 == 0),
               "tvts_gemm: pointers must be 16-byte aligned");
  TVTS_REQUIRE(g->ldo % 4 == 0, "tvts_gemm: ldo=%lld must be a multiple of 4", g->ldo);
  if (g->out_dtype == 1 || g->out_pre || g->aux)
    TVTS_REQUIRE(g->N % 8 == 0 && g->ldo % 8 == 0 && g->ldaux % 8 == 0, "tvts_gemm: bf16 outputs/aux need N, ldo, ldaux multiples of 8");
  TVTS_REQUIRE(!(g->accumulate && g->out_dtype != 0), "tvts_gemm: accumulate requires fp32 output");
  TVTS_REQUIRE(!(g->splits > 1 && !g->accumulate), "tvts_gemm: split-K requires accumulate=1 into a pre-initialised fp32 output");
  TVTS_REQUIRE(!(g->dact && !g->aux), "tvts_gemm: dact needs aux");
  TVTS_REQUIRE(!(g->splits > 1 && (g->act || g->dact || g->out_pre)), "tvts_gemm: split-K cannot be combined with act/dact/out_pre");
  TVTS_REQUIRE(!(g->out_pre && g->out_dtype != 1), "tvts_gemm: out_pre requires a bf16 output");
  if (g->residual) TVTS_REQUIRE(g->ldr % 4 == 0 && (uintptr_t)g->residual % 16 == 0, "tvts_gemm: residual alignment");
  if (g->bias) TVTS_REQUIRE((uintptr_t)g->bias % 16 == 0, "tvts_gemm: bias alignment");
  if (g->aux) TVTS_REQUIRE((uintptr_t)g->aux % 16 == 0, "tvts_gemm: aux alignment");
  if (g->out_pre) TVTS_REQUIRE((uintptr_t)g->out_pre % 16 == 0, "tvts_gemm: out_pre alignment");

  EpiParams ep;
  ep.out = g->out; ep.out_pre = g->out_pre; ep.bias = g->bias; ep.residual = g->residual;
  ep.aux = reinterpret_cast<const bf16*>(g->aux);
  ep.ldo = g->ldo; ep.ldr = g->ldr; ep.ldaux = g->ldaux;
  ep.out_dtype = g->out_dtype; ep.act = g->act; ep.dact = g->dact; ep.accumulate = g->accumulate;
  ep.alpha = g->alpha == 0.0f ? 1.0f : g->alpha;

  const bool small_n = g->N <= 128;
  const int BN = small_n ? 128 : 256;
  // CTA pairs (256 x 256 tiles) when there are enough rows to fill the machine with them ...
  bool pair = !small_n && g->M > 128;
  // ... unless wave quantisation eats their advantage: M = 25120 rows are 98.1 pair tiles, so an N = 768 problem is 297 tiles on 74
  // pair workers = 5 waves where 4.01 would do, while 128-row tiles (591 on 148 workers) fill 4 waves exactly.  A solo tile costs a CTA
  // about as long as its half of a pair tile (same MMAs per SM; ~7 % slower mainloop: each CTA stages the whole B tile), so compare
  // waves_pair against waves_solo * penalty (environment TVTS_GEMM_SOLO_PENALTY, per cent; 0 = never choose solo tiles this way).
  // OFF by default: measured on the c3 step (round 2, profiles/r2/call12_*) it speeds the N = 768 GEMMs up by 4-14 % (GEMM time 21.95 ->
  // 21.78 ms) but not the step (31.7 -> 31.9 ms): the step runs at the board's power cap, and the SM clock drops by what was gained.
  if (pair && !g->accumulate && g_solo_penalty != 0) {
    if (g_solo_penalty < 0) {
      const char* e = getenv("TVTS_GEMM_SOLO_PENALTY");
      g_solo_penalty = e ? atoi(e) : 0;
    }
    const long long nt = (g->N + BN - 1) / BN;
    const long long tiles_pair = ((g->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * nt, tiles_solo = ((g->M + BLOCK_M - 1) / BLOCK_M) * nt;
    const long long w_pair = tvts_num_sms() / 2, w_solo = tvts_num_sms();
    const long long waves_pair = (tiles_pair + w_pair - 1) / w_pair, waves_solo = (tiles_solo + w_solo - 1) / w_solo;
    if (g_solo_penalty > 0 && waves_solo * g_solo_penalty < waves_pair * 100) pair = false;
  }
  if (g_pair_mode == 0) pair = false;
  if (g_pair_mode == 1) pair = !small_n;
  const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
  GemmShape s;
  s.M = (int)g->M; s.N = (int)g->N; s.K = (int)g->K;
  s.m_tiles = (s.M + tile_m - 1) / tile_m;
  s.n_tiles = (s.N + BN - 1) / BN;
  s.k_blocks = (s.K + BLOCK_K - 1) / BLOCK_K;
  int splits = g->splits;
  if (splits <= 0) {  // auto (only when the caller accumulates): pick the split count whose units fill whole waves of workers
    splits = 1;
    if (g->accumulate) {
      const int tiles = s.m_tiles * s.n_tiles;
      const int workers = tvts_num_sms() / (pair ? 2 : 1);
      const int max_splits = (s.k_blocks + 7) / 8;  // keep >= 8 k-blocks per unit
      double best = 0.0;
      for (int cand = 1; cand <= max_splits && cand * tiles <= 3 * workers + tiles; ++cand) {
        const int units = cand * tiles;
        const int waves = (units + workers - 1) / workers;
        // efficiency of the wave packing, slightly penalising more splits (each adds a tile-sized reduce-add pass)
        const double eff = (double)units / ((double)waves * workers) - 0.004 * cand;
        if (eff > best + 1e-9) { best = eff; splits = cand; }
      }
    }
  }
  if (splits > s.k_blocks) splits = s.k_blocks;
  s.kb_per_split = (s.k_blocks + splits - 1) / splits;
  s.splits = (s.k_blocks + s.kb_per_split - 1) / s.kb_per_split;
  s.tiles = s.m_tiles * s.n_tiles;

#define TVTS_DISPATCH(BN_, C_)                                                          \
  if (!g->a_mn && !g->b_mn) return launch<BN_, false, false, C_>(g, s, ep, stream);     \
  if (g->a_mn && g->b_mn) return launch<BN_, true, true, C_>(g, s, ep, stream);         \
  if (!g->a_mn && g->b_mn) return launch<BN_, false, true, C_>(g, s, ep, stream);       \
  return launch<BN_, true, false, C_>(g, s, ep, stream);
  if (small_n) { TVTS_DISPATCH(128, 1) } else if (pair) { TVTS_DISPATCH(256, 2) } else { TVTS_DISPATCH(256, 1) }
#undef TVTS_DISPATCH
}
