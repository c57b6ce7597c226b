// TEST INFRASTRUCTURE ONLY.  A minimal host-side stand-in for the CUDA constructs the simple streaming kernels of
// tvts_b200/csrc/{v1_glue,input_stage}.cu use, so that their index arithmetic can be EXECUTED on a CPU (g++, no GPU):
// the kernel bodies are compiled unmodified (the sources include this header instead of common.cuh when TVTS_HOST_SHIM is
// defined) and tests/host_kernels/harness.cpp runs them once per (block, thread) of the launch grid.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }

struct bf16 { uint16_t bits; };   // the 16-bit operand type (bfloat16 build only)

// fp32 -> bf16, round to nearest even (what __floats2bfloat162_rn does)
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;              // NaN
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
static inline uint32_t pack_bf16x2(float lo, float hi) { return (uint32_t)f32_to_bf16_rn(lo) | ((uint32_t)f32_to_bf16_rn(hi) << 16); }
static inline float2 unpack_bf16x2(uint32_t u) {
  uint32_t a = u << 16, b = u & 0xffff0000u;
  float2 r;
  std::memcpy(&r.x, &a, 4);
  std::memcpy(&r.y, &b, 4);
  return r;
}

static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
