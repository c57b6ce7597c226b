// Fused flat AdamW for the trainer step.  Restates transformers==4.10.2 `AdamW.step` as called from
// v2/train_dist_TVTSv2_ViT_B_16.py:119-125 (un-vendored dependency; published algorithm):
//     m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; denom = sqrt(v) + eps
//     p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/denom ;  then (decoupled, AFTER the update)  p -= lr*wd * p
// over ONE flat fp32 arena holding every trainable parameter (master weights p, gradients g, moments m, v at identical
// offsets, tensors padded to whole chunks), in a single launch; the same pass refreshes the bf16 GEMM-operand copy of the
// weights.  Per-tensor hyper-parameters come from a small table indexed through a chunk -> tensor map:
//     table[t] = {step_size_t, lr_t*wd_t, active_t (0: parameter had no gradient this step -> untouched), unused}
// HBM-bound: 16 B read + 14 B written per parameter.  Compiled WITHOUT --use_fast_math (exact sqrt / division).
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, bf16* __restrict__ pb,
                                                         const int32_t* __restrict__ chunk_tensor, const float4* __restrict__ table,
                                                         int chunk_elems, float b1, float b2, float eps, float gscale) {
  const long long chunk = blockIdx.x;
  const float4 hp = __ldg(table + __ldg(chunk_tensor + chunk));
  if (hp.z == 0.0f) return;  // block-uniform
  const float step_size = hp.x, lrwd = hp.y;
  const long long base = chunk * (long long)chunk_elems;
  for (int i = threadIdx.x * 4; i < chunk_elems; i += 256 * 4) {
    const long long o = base + i;
    float4 pp = *reinterpret_cast<const float4*>(p + o);
    const float4 gg = *reinterpret_cast<const float4*>(g + o);
    float4 mm = *reinterpret_cast<const float4*>(m + o);
    float4 vv = *reinterpret_cast<const float4*>(v + o);
    float* P = reinterpret_cast<float*>(&pp);
    const float* G = reinterpret_cast<const float*>(&gg);
    float* Mo = reinterpret_cast<float*>(&mm);
    float* Vo = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = G[j] * gscale;
      Mo[j] = Mo[j] * b1 + (1.0f - b1) * gj;
      Vo[j] = Vo[j] * b2 + (1.0f - b2) * gj * gj;
      const float denom = sqrtf(Vo[j]) + eps;
      float x = P[j] - step_size * (Mo[j] / denom);
      if (lrwd > 0.0f) x = x - lrwd * x;
      P[j] = x;
    }
    *reinterpret_cast<float4*>(p + o) = pp;
    *reinterpret_cast<float4*>(m + o) = mm;
    *reinterpret_cast<float4*>(v + o) = vv;
    if (pb != nullptr) *reinterpret_cast<uint2*>(pb + o) = make_uint2(pack_bf16x2(P[0], P[1]), pack_bf16x2(P[2], P[3]));
  }
}

// ------------------------------------------------------------------------------------------------ dynamic loss scaling
// 16-bit IEEE-half operands (TVTS_OPERAND=fp16) run the backward under a loss scale.  Everything that decides whether a step counts
// lives on the device, so the whole step stays one CUDA graph:
//   state[0] loss scale S          state[1] consecutive finite steps       state[2] non-finite gradient seen this step (0/1)
//   state[3] steps skipped so far
//   steps[t] optimizer step count of tensor t (bias correction is evaluated in the kernel from it)
// launch order per step: grad_check -> adamw_flat_dyn (no-op when state[2] is set) -> scale_update (back-off x0.5 and skip, or count
// the step and grow x2 every `growth_interval` finite steps: torch.cuda.amp.GradScaler's policy)
__global__ void __launch_bounds__(256) grad_check_kernel(const float4* __restrict__ g, long long n4, float* __restrict__ state) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(g + i);
    const uint32_t e = (__float_as_uint(x.x) & 0x7f800000u) == 0x7f800000u || (__float_as_uint(x.y) & 0x7f800000u) == 0x7f800000u ||
                       (__float_as_uint(x.z) & 0x7f800000u) == 0x7f800000u || (__float_as_uint(x.w) & 0x7f800000u) == 0x7f800000u;
    bad |= e != 0;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) state[2] = 1.0f;
}

__global__ void __launch_bounds__(256) adamw_flat_dyn_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                             float* __restrict__ v, bf16* __restrict__ pb,
                                                             const int32_t* __restrict__ chunk_tensor, const float4* __restrict__ table,
                                                             const int32_t* __restrict__ steps, const float* __restrict__ state,
                                                             int chunk_elems, float b1, float b2, float eps) {
  if (state[2] != 0.0f) return;   // a non-finite gradient somewhere: the whole step is skipped (grid-uniform)
  const long long chunk = blockIdx.x;
  const int tensor = __ldg(chunk_tensor + chunk);
  const float4 hp = __ldg(table + tensor);   // {lr, lr*wd, active, correct_bias}
  if (hp.z == 0.0f) return;
  const float t = (float)(__ldg(steps + tensor) + 1);
  const float step_size = hp.w != 0.0f ? hp.x * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t)) : hp.x;
  const float lrwd = hp.y;
  const float gscale = 1.0f / state[0];
  const long long base = chunk * (long long)chunk_elems;
  for (int i = threadIdx.x * 4; i < chunk_elems; i += 256 * 4) {
    const long long o = base + i;
    float4 pp = *reinterpret_cast<const float4*>(p + o);
    const float4 gg = *reinterpret_cast<const float4*>(g + o);
    float4 mm = *reinterpret_cast<const float4*>(m + o);
    float4 vv = *reinterpret_cast<const float4*>(v + o);
    float* P = reinterpret_cast<float*>(&pp);
    const float* G = reinterpret_cast<const float*>(&gg);
    float* Mo = reinterpret_cast<float*>(&mm);
    float* Vo = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = G[j] * gscale;
      Mo[j] = Mo[j] * b1 + (1.0f - b1) * gj;
      Vo[j] = Vo[j] * b2 + (1.0f - b2) * gj * gj;
      const float denom = sqrtf(Vo[j]) + eps;
      float x = P[j] - step_size * (Mo[j] / denom);
      if (lrwd > 0.0f) x = x - lrwd * x;
      P[j] = x;
    }
    *reinterpret_cast<float4*>(p + o) = pp;
    *reinterpret_cast<float4*>(m + o) = mm;
    *reinterpret_cast<float4*>(v + o) = vv;
    if (pb != nullptr) *reinterpret_cast<uint2*>(pb + o) = make_uint2(pack_bf16x2(P[0], P[1]), pack_bf16x2(P[2], P[3]));
  }
}

__global__ void scale_update_kernel(float* __restrict__ state, int32_t* __restrict__ steps, const float4* __restrict__ table, int n_tensors,
                                    float growth_interval, float max_scale) {
  const bool bad = state[2] != 0.0f;
  __syncthreads();
  if (!bad)
    for (int t = threadIdx.x; t < n_tensors; t += blockDim.x)
      if (table[t].z != 0.0f) steps[t] += 1;
  if (threadIdx.x == 0) {
    if (bad) {
      state[0] = fmaxf(state[0] * 0.5f, 1.0f);
      state[1] = 0.0f;
      state[3] += 1.0f;
    } else {
      state[1] += 1.0f;
      if (state[1] >= growth_interval) { state[0] = fminf(state[0] * 2.0f, max_scale); state[1] = 0.0f; }
    }
    state[2] = 0.0f;
  }
}

}  // namespace

extern "C" int tvts_adamw_dyn_check(const float* g, int64_t n_elems, float* state, void* stream) {
  if (n_elems == 0) return TVTS_OK;
  TVTS_REQUIRE(g && state, "adamw_dyn_check: null pointer");
  TVTS_REQUIRE(n_elems > 0 && n_elems % 4 == 0 && (uintptr_t)g % 16 == 0, "adamw_dyn_check: n_elems=%lld must be a multiple of 4, g 16-byte aligned",
               (long long)n_elems);
  const long long n4 = n_elems / 4;
  const int blocks = (int)((n4 + 255) / 256 < (long long)tvts_num_sms() * 8 ? (n4 + 255) / 256 : (long long)tvts_num_sms() * 8);
  grad_check_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(g), n4, state);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_adamw_dyn_apply(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                                    const int32_t* steps, const float* state, int64_t n_chunks, int64_t chunk_elems, float beta1, float beta2,
                                    float eps, void* stream) {
  if (n_chunks == 0) return TVTS_OK;
  TVTS_REQUIRE(p && g && m && v && chunk_tensor && table && steps && state, "adamw_dyn_apply: null pointer");
  TVTS_REQUIRE(chunk_elems > 0 && chunk_elems % 4 == 0, "adamw_dyn_apply: chunk_elems=%lld must be a positive multiple of 4", (long long)chunk_elems);
  TVTS_REQUIRE(n_chunks > 0 && n_chunks < (1ll << 31), "adamw_dyn_apply: bad sizes");
  TVTS_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)table) % 16 == 0 && (uintptr_t)p_bf16 % 8 == 0,
               "adamw_dyn_apply: arena pointers must be 16-byte aligned");
  adamw_flat_dyn_kernel<<<(unsigned)n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, reinterpret_cast<bf16*>(p_bf16), chunk_tensor, reinterpret_cast<const float4*>(table), steps, state, (int)chunk_elems, beta1,
      beta2, eps);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_adamw_dyn_finish(int32_t* steps, const float* table, float* state, int64_t n_tensors, float growth_interval,
                                     float max_scale, void* stream) {
  TVTS_REQUIRE(steps && table && state, "adamw_dyn_finish: null pointer");
  TVTS_REQUIRE(n_tensors > 0 && n_tensors < (1ll << 31) && (uintptr_t)table % 16 == 0, "adamw_dyn_finish: bad sizes");
  scale_update_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state, steps, reinterpret_cast<const float4*>(table), (int)n_tensors,
                                                                           growth_interval, max_scale);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_adamw_flat_dyn(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                                   int32_t* steps, float* state, int64_t n_tensors, int64_t n_chunks, int64_t chunk_elems, float beta1,
                                   float beta2, float eps, float growth_interval, float max_scale, void* stream) {
  if (n_chunks == 0) return TVTS_OK;
  TVTS_REQUIRE(p && g && m && v && chunk_tensor && table && steps && state, "adamw_flat_dyn: null pointer");
  TVTS_REQUIRE(chunk_elems > 0 && chunk_elems % 4 == 0, "adamw_flat_dyn: chunk_elems=%lld must be a positive multiple of 4", (long long)chunk_elems);
  TVTS_REQUIRE(n_chunks < (1ll << 31) && n_tensors > 0 && n_tensors < (1ll << 31), "adamw_flat_dyn: bad sizes");
  int rc = tvts_adamw_dyn_check(g, n_chunks * chunk_elems, state, stream);
  if (rc) return rc;
  rc = tvts_adamw_dyn_apply(p, g, m, v, p_bf16, chunk_tensor, table, steps, state, n_chunks, chunk_elems, beta1, beta2, eps, stream);
  if (rc) return rc;
  return tvts_adamw_dyn_finish(steps, table, state, n_tensors, growth_interval, max_scale, stream);
}

extern "C" int tvts_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                               int64_t n_chunks, int64_t chunk_elems, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (n_chunks == 0) return TVTS_OK;
  TVTS_REQUIRE(p && g && m && v && chunk_tensor && table, "adamw_flat: null pointer");
  TVTS_REQUIRE(chunk_elems > 0 && chunk_elems % 4 == 0, "adamw_flat: chunk_elems=%lld must be a positive multiple of 4", (long long)chunk_elems);
  TVTS_REQUIRE(n_chunks < (1ll << 31), "adamw_flat: too many chunks");
  TVTS_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)table) % 16 == 0 && (uintptr_t)p_bf16 % 8 == 0,
               "adamw_flat: arena pointers must be 16-byte aligned");
  adamw_flat_kernel<<<(unsigned)n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, reinterpret_cast<bf16*>(p_bf16), chunk_tensor, reinterpret_cast<const float4*>(table), (int)chunk_elems, beta1, beta2, eps,
      grad_scale);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
