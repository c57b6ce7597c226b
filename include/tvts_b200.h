/* tvts_b200 -- C ABI of the B200-native TVTS/TVTSv2 pre-training hot path.
 *
 * The reference (TencentARC/TVTS) has no native/FFI layer: its hot path is stock PyTorch modules
 * (SURVEY.md section 8b).  This header is therefore the boundary our own Python mirror of the reference
 * modules (tvts_b200/v2/model/..., tvts_b200/v2/trainer/...) binds through ctypes; each entry point names
 * the reference computation (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - caller-allocated buffers, no ownership transfer, no hidden allocation
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no internal sync
 *   - return 0 on success, negative on error; tvts_last_error() returns the thread-local message
 *   - bf16 = IEEE bfloat16 stored as uint16_t; "f32" = float
 */
#ifndef TVTS_B200_H
#define TVTS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVTS_B200_VERSION 100

int tvts_version(void);
const char* tvts_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
long long tvts_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM:  out[M,N] = epilogue(alpha * sum_k A[m,k]*B[n,k])      (tcgen05 + TMA, bf16 in / fp32 accumulate)
 * replaces every nn.Linear / `x @ proj` on the path:
 *   v2/model/video_encoder_ViT_B_16.py:41 (qkv), :74 (proj), :105-109 (mlp), :180 (conv1 as GEMM), :233 (x @ proj)
 *   v2/CLIP/clip/model.py:175-192 (text blocks), v2/model/sort_transformer.py:29-33,43-57 (sort head)
 * and their autograd backward (dgrad, wgrad).
 *   a_mn / b_mn = 0: operand stored [rows, K] row-major (ld = elements per row)      "K-major"
 *               = 1: operand stored [K, rows] row-major                               "MN-major" (wgrad)
 * epilogue order: v = alpha*acc + bias[n]; out_pre = bf16(v) (optional); v = act(v); v *= act'(aux[m,n]) (dact);
 *                 v += residual[m,n]; out = v (store, or atomic add when accumulate=1)
 * splits: 0 = auto (split-K only when accumulate=1), >1 requires accumulate=1.
 */
typedef struct tvts_gemm_args {
  const void* a;        /* bf16 */
  const void* b;        /* bf16 */
  void* out;            /* f32 or bf16 [M, ldo] */
  void* out_pre;        /* optional bf16 [M, ldo]: value before the activation */
  const float* bias;    /* optional f32 [N] */
  const float* residual;/* optional f32 [M, ldr] */
  const void* aux;      /* optional bf16 [M, ldaux]: pre-activation for dact */
  int64_t M, N, K;
  int64_t lda, ldb, ldo, ldr, ldaux;
  int32_t a_mn, b_mn;
  int32_t out_dtype;    /* 0 = f32, 1 = bf16 */
  int32_t act;          /* 0 none, 1 QuickGELU, 2 GELU(erf) */
  int32_t dact;         /* 0 none, else multiply by derivative of that activation at aux */
  int32_t accumulate;   /* 1: out += result (fp32 atomics) */
  int32_t splits;
  float alpha;          /* 0 is treated as 1 */
} tvts_gemm_args;
int tvts_gemm(const tvts_gemm_args* args, void* stream);
/* debug knob for bring-up of the MN-major shared-memory descriptors (0 = built-in defaults) */
int tvts_gemm_debug_set(int lbo_bytes, int sbo_bytes, int k_advance_bytes);
/* debug knob: 0 normal epilogue, 1 no global stores, 2 direct row-per-thread bf16 stores, 3 skip epilogue (timing experiments) */
int tvts_gemm_debug_epi(int mode);

#ifdef __cplusplus
}
#endif
#endif /* TVTS_B200_H */
