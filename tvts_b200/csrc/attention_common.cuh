// Head-dim independent pieces of the attention kernels (attention.cu: head dim 64 fast paths; attention_hd.cu: generic head dim):
// launch geometry, the index sets of the three attention modes, and the cp.async / ldmatrix / mma.sync primitives.
#pragma once
#ifdef TVTS_HOST_SHIM          // tests/host_kernels: the kernels are also compiled for a CPU SIMT stand-in that supplies the primitives below
#include "host_simt.h"
#else
#include "common.cuh"
#endif

namespace {

constexpr int BM = 64;          // stationary rows per CTA (streamed kernels): 4 warps x 16
constexpr int BN = 64;          // streamed rows per tile
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnShape {
  int B, N, H;
  int mode;    // 0 full, 1 space, 2 time
  int T, n;    // frames, kept tokens per frame (modes 1, 2)
  int causal;  // mode 0 only
  float scale;
  int cls_only;  // modes 1/2: launch covers only the CLS group (used next to the time kernels)
  int q0, qn;    // mode 0 only: QUERY WINDOW [q0, q0+qn) (qn = 0: all N rows are queries); keys are always all N tokens
};

struct Sets {
  int st_base, st_stride, st_count;
  int sm_has0, sm_base, sm_stride, sm_count;
};

__host__ __device__ inline int chunks_per_group(const AttnShape& a) {
  if (a.mode == 0) return (a.N + BM - 1) / BM;
  const int len = a.mode == 1 ? a.n : a.T;
  return (len + BM - 1) / BM;
}
// grid.x of a streamed launch; `transposed` = the dK/dV role (stationary rows are keys)
__host__ __device__ inline int num_blocks_x(const AttnShape& a, bool transposed = false) {
  if (a.mode == 0) return (a.qn > 0 && !transposed) ? (a.qn + BM - 1) / BM : chunks_per_group(a);
  if (a.cls_only) return 1;
  const int groups = a.mode == 1 ? a.T : a.n;
  return groups * chunks_per_group(a) + 1;  // + the CLS group
}

__device__ inline Sets decode_sets(const AttnShape& a, int bx, bool transposed = false) {
  Sets s;
  if (a.mode == 0) {
    if (a.qn > 0 && !transposed) {        // stationary = the query window, streamed = every key
      s.st_base = a.q0 + bx * BM; s.st_stride = 1; s.st_count = min(BM, a.q0 + a.qn - s.st_base);
      s.sm_has0 = 0; s.sm_base = 0; s.sm_stride = 1; s.sm_count = a.N;
    } else if (a.qn > 0) {                // stationary = every key, streamed = the query window
      s.st_base = bx * BM; s.st_stride = 1; s.st_count = min(BM, a.N - s.st_base);
      s.sm_has0 = 0; s.sm_base = a.q0; s.sm_stride = 1; s.sm_count = a.qn;
    } else {
      s.st_base = bx * BM; s.st_stride = 1; s.st_count = min(BM, a.N - s.st_base);
      s.sm_has0 = 0; s.sm_base = 0; s.sm_stride = 1; s.sm_count = a.N;
    }
    return s;
  }
  const int cpg = chunks_per_group(a);
  const int groups = a.mode == 1 ? a.T : a.n;
  const int g = a.cls_only ? groups : bx / cpg, c = a.cls_only ? 0 : bx - g * cpg;
  if (g >= groups) {  // CLS token <-> all tokens
    s.st_base = 0; s.st_stride = 1; s.st_count = 1;
    s.sm_has0 = 0; s.sm_base = 0; s.sm_stride = 1; s.sm_count = a.N;
    return s;
  }
  if (a.mode == 1) {
    s.st_base = 1 + g * a.n + c * BM; s.st_stride = 1; s.st_count = min(BM, a.n - c * BM);
    s.sm_has0 = 1; s.sm_base = 1 + g * a.n; s.sm_stride = 1; s.sm_count = a.n;
  } else {
    s.st_base = 1 + g + c * BM * a.n; s.st_stride = a.n; s.st_count = min(BM, a.T - c * BM);
    s.sm_has0 = 1; s.sm_base = 1 + g; s.sm_stride = a.n; s.sm_count = a.T;
  }
  return s;
}

__device__ __forceinline__ int streamed_token(const Sets& s, int k) {
  return (s.sm_has0 && k == 0) ? 0 : s.sm_base + (k - s.sm_has0) * s.sm_stride;
}

// ------------------------------------------------------------------------------------------------ primitives
#ifndef TVTS_HOST_SHIM

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// D(16x8, fp32) += A(16x16 bf16, row) * B(16x8 bf16, col)
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." TVTS_MMA_TYPE "." TVTS_MMA_TYPE ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#endif  // !TVTS_HOST_SHIM

}  // namespace
