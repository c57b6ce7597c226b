// TEST INFRASTRUCTURE ONLY.  A small SIMT stand-in on top of host_shim.h that lets the warp-level tensor-core kernels of
// tvts_b200/csrc/attention_hd.cu be EXECUTED on a CPU: every CUDA thread of a CTA is an OS thread, __syncthreads / __syncwarp are
// barriers, and the five PTX primitives the kernels use (cp.async 16-byte copies, ldmatrix x4 / x4.trans, mma.sync m16n8k16 bf16,
// shfl.xor) are restated functionally with their architectural lane <-> element layouts.  The same primitives with the same layouts
// drive the GPU-verified head-dim-64 kernels of attention.cu, so a layout mistake here shows up as a mismatch at HD = 64 as well.
#pragma once
#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#include "host_shim.h"

#define __shared__ static                 // a function-local static array is shared by all (OS-thread) CUDA threads: one CTA runs at a time
#define __align__(x) __attribute__((aligned(x)))
#define TVTS_DYN_SMEM(type, name, align) type* const name = reinterpret_cast<type*>(simt::dyn_smem)
#define __expf(x) expf(x)
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#ifndef INFINITY
#define INFINITY (__builtin_inff())
#endif

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline bf16 opnd_from_float(float x) { return bf16{f32_to_bf16_rn(x)}; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#define __logf(x) logf(x)      // (glibc declares a __logf of its own)

namespace simt {
struct Warp {
  std::barrier<> bar{32};
  uint32_t u32[32][6];
  uint32_t addr[32];
  float f32[32];
};
alignas(128) inline uint8_t dyn_smem[232 * 1024];      // the dynamic shared memory of the running CTA
// anchor of the 32-bit "shared window" addresses: every static array of this module (the kernels' `__shared__` arrays included) lies
// within +-2 GB of it.  Kept as an integer so that no object bounds are attached to the derived pointers.
inline uintptr_t smem_anchor = reinterpret_cast<uintptr_t>(dyn_smem);
inline std::vector<Warp>* warps = nullptr;
inline std::barrier<>* cta_bar = nullptr;
inline thread_local int lane = 0, warp = 0;
inline Warp& W() { return (*warps)[warp]; }

// run `body` once per CUDA thread of every CTA of the grid; CTAs run one after the other, the threads of a CTA concurrently
inline void launch(unsigned gx, unsigned gy, unsigned gz, unsigned block, const std::function<void()>& body) {
  for (unsigned bz = 0; bz < gz; ++bz)
    for (unsigned by = 0; by < gy; ++by)
      for (unsigned bx = 0; bx < gx; ++bx) {
        std::vector<Warp> w((block + 31) / 32);
        std::barrier<> cb((std::ptrdiff_t)block);
        warps = &w;
        cta_bar = &cb;
        std::vector<std::thread> th;
        for (unsigned tx = 0; tx < block; ++tx)
          th.emplace_back([=, &body] {
            gridDim.x = gx; gridDim.y = gy; gridDim.z = gz; blockDim.x = block;
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz; threadIdx.x = tx;
            lane = (int)(tx & 31); warp = (int)(tx >> 5);
            body();
          });
        for (auto& t : th) t.join();
      }
}
}  // namespace simt

static inline void __syncthreads() { simt::cta_bar->arrive_and_wait(); }
static inline void __syncwarp() { simt::W().bar.arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lanemask) {
  auto& w = simt::W();
  w.f32[simt::lane] = v;
  w.bar.arrive_and_wait();
  const float r = w.f32[simt::lane ^ lanemask];
  w.bar.arrive_and_wait();
  return r;
}
static inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
static inline uint32_t smem_u32(const void* p) { return (uint32_t)(int32_t)(int64_t)(reinterpret_cast<uintptr_t>(p) - simt::smem_anchor); }
static inline uint8_t* smem_ptr(uint32_t a) { return reinterpret_cast<uint8_t*>(simt::smem_anchor + (uintptr_t)(int64_t)(int32_t)a); }

// cp.async.cg.shared.global 16 bytes (src-size 0 -> zero fill); executed synchronously, so commit / wait are no-ops
static inline void cp_async16(uint32_t dst, const void* src, bool valid) {
  if (valid) std::memcpy(smem_ptr(dst), src, 16);
  else std::memset(smem_ptr(dst), 0, 16);
}
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}

// ldmatrix.sync.aligned.m8n8.x4(.trans).shared.b16: row r of matrix i is the 16 bytes at the address supplied by lane 8i + r;
// plain: lane T receives M[T/4][2(T%4)], M[T/4][2(T%4)+1];  .trans: lane T receives M[2(T%4)][T/4], M[2(T%4)+1][T/4]
static inline void ldsm_impl(uint32_t addr, uint32_t* r, bool trans) {
  auto& w = simt::W();
  const int T = simt::lane;
  w.addr[T] = addr;
  w.bar.arrive_and_wait();
  for (int i = 0; i < 4; ++i) {
    if (!trans) {
      std::memcpy(&r[i], smem_ptr(w.addr[8 * i + T / 4]) + 4 * (T % 4), 4);
    } else {
      uint16_t lo, hi;
      std::memcpy(&lo, smem_ptr(w.addr[8 * i + 2 * (T % 4)]) + 2 * (T / 4), 2);
      std::memcpy(&hi, smem_ptr(w.addr[8 * i + 2 * (T % 4) + 1]) + 2 * (T / 4), 2);
      r[i] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
  }
  w.bar.arrive_and_wait();
}
static inline void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  uint32_t r[4];
  ldsm_impl(addr, r, false);
  r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}
static inline void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  uint32_t r[4];
  ldsm_impl(addr, r, true);
  r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}

static inline float bf16_half(uint32_t word, int half) {
  const float2 f = unpack_bf16x2(word);
  return half ? f.y : f.x;
}
// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32, fragment layouts (g = lane/4, t = lane%4):
//   A: a0 = (row g,   k 2t,2t+1)  a1 = (row g+8, k 2t,2t+1)  a2 = (row g,   k 2t+8,2t+9)  a3 = (row g+8, k 2t+8,2t+9)
//   B: b0 = (k 2t,2t+1; col g)    b1 = (k 2t+8,2t+9; col g)
//   C/D: c0,c1 = (row g; cols 2t,2t+1)   c2,c3 = (row g+8; cols 2t,2t+1)
static inline void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  auto& w = simt::W();
  const int T = simt::lane, g = T / 4, t = T % 4;
  for (int i = 0; i < 4; ++i) w.u32[T][i] = a[i];
  w.u32[T][4] = b0;
  w.u32[T][5] = b1;
  w.bar.arrive_and_wait();
  for (int ri = 0; ri < 2; ++ri)
    for (int ci = 0; ci < 2; ++ci) {
      const int row = g + 8 * ri, col = 2 * t + ci;
      float s = 0.f;
      for (int k = 0; k < 16; ++k) {
        const float av = bf16_half(w.u32[(row % 8) * 4 + (k % 8) / 2][(row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)], k % 2);
        const float bv = bf16_half(w.u32[col * 4 + (k % 8) / 2][4 + (k >= 8 ? 1 : 0)], k % 2);
        s += av * bv;
      }
      c[2 * ri + ci] += s;
    }
  w.bar.arrive_and_wait();
}
