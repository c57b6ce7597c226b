/* tvts_b200 -- C ABI of the B200-native TVTS/TVTSv2 pre-training hot path.
 *
 * The reference (TencentARC/TVTS) has no native/FFI layer: its hot path is stock PyTorch modules
 * (SURVEY.md section 8b).  This header is therefore the boundary our own Python mirror of the reference
 * modules (tvts_b200/modules.py, modules_v1.py, trainer.py; drop-in trees tvts_b200/dropin, dropin_v1) binds through ctypes
 * (tvts_b200/_lib.py); each entry point names
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
/* 16-bit operand format this library was built for: 0 = bfloat16 (libtvts_b200.so, the default), 1 = IEEE half (libtvts_b200_fp16.so,
 * built with -DTVTS_OPERAND_FP16).  Everything this header calls "bf16" (GEMM operands, 16-bit activations and gradients, the weight
 * shadow arena) is in that format; fp32 tensors are unaffected. */
int tvts_operand_format(void);
const char* tvts_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
long long tvts_launch_count(void);

/* live timing of the GEMM launches (bench.py roofline): enable(1) resets and starts recording one CUDA-event pair per
 * tvts_gemm launch on its stream; after synchronising, collect() returns the number of launches recorded and their summed
 * device time (ms), algorithmic FLOPs (2*M*N*K) and bytes. */
int tvts_prof_enable(int on);
long long tvts_prof_collect(double* total_ms, double* total_flops, double* total_bytes);
/* per-launch readout before collect(): tags = {M, N, K, flags: 1 a_mn, 2 b_mn, 4 CTA pair, 8 bf16 out, 16 residual, 32 out_pre, 64 dact, 128 accumulate, splits<<8} */
int tvts_prof_count(void);
int tvts_prof_record(int i, double* ms, double* flops, long long* tags);

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
/* tile policy: -1 auto (CTA pairs / cta_group::2 with 256x256 tiles for large problems), 0 single-CTA 128x256 tiles only, 1 pairs always */
int tvts_gemm_set_pair_mode(int mode);
/* 1 (default): the residual / aux epilogue operand is prefetched by TMA into the store ring; 0: read straight from global memory */
int tvts_gemm_set_operand_prefetch(int on);
/* 0 (default): results leave through 128B-swizzled shared-memory boxes and TMA stores; 1: results that are STORED (not accumulated:
 * accumulate = 1 always uses cp.reduce.async.bulk .add) leave straight from registers with 32-byte stores (measured neutral on the
 * hot path; environment TVTS_GEMM_EPI_DIRECT=1 selects it for a whole process) */
int tvts_gemm_set_epilogue_direct(int on);
/* debug: co-resident clusters of the pair kernel for a given cluster size (cudaOccupancyMaxActiveClusters) */
int tvts_gemm_debug_max_clusters(int cluster_size);
/* debug knob: 0 normal epilogue, 1 no global stores, 2 direct row-per-thread bf16 stores, 3 skip epilogue (timing experiments) */
int tvts_gemm_debug_epi(int mode);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (fp32 statistics, one warp per row).  Replaces LayerNorm.forward of
 * v2/model/video_encoder_ViT_B_16.py:79-85, v2/CLIP/clip/model.py:157-163, nn.LayerNorm(eps=1e-6) of
 * v2/model/sort_transformer.py:73,76,100 and their backward.  D must be a multiple of 128 (<= 1280).
 *   fwd: y = (x-mean)*rstd*gamma+beta as bf16 (GEMM operand) or f32; mean/rstd [M] saved for backward
 *   bwd: dx = LN'(dy) + res1 + res2 (both optional) as f32 and/or bf16; dgamma/dbeta [D] are ACCUMULATED (atomicAdd)
 */
int tvts_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int64_t y_is_bf16, float* mean, float* rstd,
                       int64_t M, int64_t D, float eps, void* stream);
int tvts_layernorm_bwd(const void* dy, int64_t dy_is_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* res1, const float* res2, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int64_t M,
                       int64_t D, void* stream);

/* layernorm_bwd that also ACCUMULATES dxsum[D] += column sums of dx: dx is the output gradient of the Linear that produced this
 * LayerNorm's input branch (attn.proj / mlp.c_proj), so its column sums are that Linear's bias gradient -- no separate pass. */
int tvts_layernorm_bwd_colsum(const void* dy, int64_t dy_is_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                              const float* res1, const float* res2, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum,
                              int64_t M, int64_t D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Grouped multi-head attention on the packed qkv buffer [B, N, 3, H, d] (bf16) written by the qkv GEMM; head dim d = 64 (B/16, B/32,
 * text towers, sort head: specialised kernels) or 80 (ViT-H/14 video tower, v2/model/model_dist_TVTSv2_ViT_H_14.py:43-45: generic kernels).
 *   mode 0 FULL  (+causal): v2/CLIP/clip/model.py:185-188 (nn.MultiheadAttention + mask :330-336),
 *                           v2/model/sort_transformer.py:9-13,43-57
 *   mode 1 SPACE / mode 2 TIME: VarAttention, v2/model/video_encoder_ViT_B_16.py:38-76 (N = 1 + T*n, token 0 = CLS)
 * out [B, N, H, d] bf16; lse [B, H, N] f32 (saved for backward); scale multiplies q.k (reference: q * d^-0.5, :45).
 * bwd writes every element of dqkv (same layout as qkv); delta_ws is a [B, H, N] f32 workspace.
 */
int tvts_attn_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n,
                  int64_t causal, float scale, void* stream);
/* 1 (default): the CLS row/column launches of the divided modes run on an internal side stream forked from / joined into `stream`
 * with events (graph-capture safe); 0: everything on `stream` */
int tvts_attn_set_side_stream(int on);
/* tcgen05 path of tvts_attn_fwd / tvts_attn_bwd (csrc/attention_tc.cu) for head dim 64 and groups that fit ONE 128-row UMMA tile:
 * mode 1 (space; n + 1 <= 128 rows per frame group, the CLS query / key folded into every frame tile and merged inside the kernel)
 * mode 0 with N <= 128 (the 77-token causal CLIP text sequences), and mode 2 (time: a tile packs floor(127 / T) patch positions x T
 * frames + CLS, block-diagonal key mask, 4-D TMA boxes over the strided token rows).  S and O (backward: S, dP, dV, dK, dQ) live in TMEM,
 * Q / K / V / dO tiles arrive by TMA; the forward's result leaves by a TMA store, the backward (a persistent kernel that draws its tiles
 * from a dynamic scheduler) stores dq / dk / dv straight from registers.  tvts_attn_set_tc(on): bit 0 = modes 0 / 1, bit 1 = mode 2 (default 3;
 * 0 routes every shape back to the mma.sync kernels -- A/B comparison in the tests; environment TVTS_ATTN_TC=0 / TVTS_ATTN_TC_TIME=0 do
 * the same per process); tvts_attn_tc_supported tells which path tvts_attn_fwd / _bwd take for a shape.
 * The CLS partials, the per-column merge tickets (modes 1 / 2) and the backward's tile-scheduler state live in ONE per-device workspace
 * (allocated on first use, outside any stream capture).  Calls of the same KIND on one device -- forward or backward, mode 0 or modes
 * 1 / 2 -- must therefore be ordered with respect to each other (one stream, or explicit dependencies); different kinds may overlap
 * (the text tower's mode-0 calls run on a second stream next to the video tower's mode-1 / 2 calls). */
int tvts_attn_set_tc(int on);
int tvts_attn_tc_supported(int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal);
int tvts_attn_tc_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n,
                     int64_t causal, float scale, void* stream);
int tvts_attn_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int64_t B, int64_t N, int64_t H,
                     int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale, void* stream);
/* the same, additionally ACCUMULATING the bias gradient of the qkv Linear (v2/model/video_encoder_ViT_B_16.py:26,41) into dbias[3*H*d]
 * (fp32) inside the kernel instead of a separate pass over dqkv -- in exact-arithmetic form, see tvts_attn_bwd_bias */
int tvts_attn_tc_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias, int64_t B, int64_t N,
                          int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale, void* stream);
int tvts_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int64_t B, int64_t N,
                  int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale, void* stream);

/* tvts_attn_bwd that also ACCUMULATES the qkv Linear's bias gradient, dbias[3*H*d] += column sums of dqkv over the tokens (fp32).  The
 * tcgen05 kernels produce it inside the backward, in exact-arithmetic form: q third = column sums of dQ (from the fp32 accumulators);
 * v third = column sums of dO (every softmax row sums to 1, so sum_j dV[j,:] = sum_i dO[i,:]); k third = 0 (a constant added to every key
 * shifts all scores of a query alike: its gradient vanishes identically, what a column sum of the stored dK only approximates by its
 * rounding noise).  Elsewhere: tvts_attn_bwd + tvts_colsum_bf16 (column sums of the stored 16-bit dqkv). */
int tvts_attn_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, float* dbias,
                       int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale,
                       void* stream);
/* The head-dim generic streamed kernels behind tvts_attn_fwd / tvts_attn_bwd for d != 64, callable directly (d = 64 or 80; same
 * arguments and results): with d = 64 they cross-check the generic code against the specialised kernels. */
/* 1 (default): groups that fit one CTA (space attention of H/14: 77 rows; short full-attention sequences) use the group-resident generic
 * kernels + a CLS launch; 0: streamed generic kernels only */
int tvts_attn_hd_set_group(int on);
/* 1 (default): with the fast paths, the CLS row / column launch of the divided modes runs on an internal side stream (fork / join with
 * events, capture-safe), like tvts_attn_set_side_stream for the head-dim-64 kernels; 0: everything on `stream` */
int tvts_attn_hd_set_side_stream(int on);
int tvts_attn_generic_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T,
                          int64_t n, int64_t causal, float scale, void* stream);
int tvts_attn_generic_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int64_t B,
                          int64_t N, int64_t H, int64_t d, int64_t mode, int64_t T, int64_t n, int64_t causal, float scale, void* stream);

/* Key-padded full attention for the TVTS v1 text encoder (DistilBERT, an un-vendored `transformers` dependency; call site
 * v1/model/model_dist_TVTS.py:124-126): keys at positions >= klen[b] (the right-padded tail of `attention_mask`) are excluded from the
 * softmax; query rows past klen[b] are still computed, like the reference.  klen [B] int32, 1 <= klen[b] <= N.  d = 64 or 80. */
int tvts_attn_padded_fwd(const void* qkv, void* out, float* lse, const int32_t* klen, int64_t B, int64_t N, int64_t H, int64_t d,
                         float scale, void* stream);
int tvts_attn_padded_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                         const int32_t* klen, int64_t B, int64_t N, int64_t H, int64_t d, float scale, void* stream);

/* Query-window attention (mode FULL, non-causal): only rows [q0, q0+qn) of every sample are queries, all N tokens are keys/values.
 * Used for the LAST block of the sort head, whose only consumed outputs are its n_trans transcript rows
 * (v2/model/sort_transformer.py:134-142).  fwd writes out / lse for the window rows only.  bwd writes dq for the window rows only
 * (the caller zero-fills dqkv first) and dk / dv for every token; rows of out / dout outside the window must be finite (zero). */
int tvts_attn_window_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t N, int64_t H, int64_t d, int64_t q0, int64_t qn,
                         float scale, void* stream);
int tvts_attn_window_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int64_t B,
                         int64_t N, int64_t H, int64_t d, int64_t q0, int64_t qn, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Video token front end (v2/model/video_encoder_ViT_B_16.py:176-216).
 *   patch_gather: video [B,T,3,R,R] f32 + keep_ind [B,n] int64 -> im2col rows of the KEPT patches only,
 *                 cols [(b*T+t)*n+j, c*p*p+u*p+v] bf16 (Conv2d k=s=p is a per-patch linear map, so the tube mask
 *                 :200-216 is applied before the patch-embed GEMM)
 *   video_assemble: x0[b,0] = cls + pos[0]; x0[b,1+t*n+j] = tok[(b*T+t)*n+j] + pos[1+keep[b,j]] + tem[t]
 *   video_assemble_bwd: dcls/dpos/dtem ACCUMULATED (atomicAdd), dtok written as bf16 (operand of the conv1 wgrad GEMM)
 */
int tvts_patch_gather(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p, int64_t n,
                      void* stream);
/* Patch sizes that are not a multiple of 4 (ViT-H/14: 14x14 patches, v2/model/video_encoder_ViT_H_14.py:419-441): 3*p*p = 588 bf16 is
 * not a whole number of 16-byte units, so the im2col rows are written at a pitch of `ld` elements (a multiple of 8, >= 3*p*p) with a
 * zero tail, and the conv1 weight is cast to bf16 rows of the same pitch (cast_bf16_pad: src [rows, cols] f32 contiguous ->
 * dst [rows, ld] bf16, tail zero): the patch-embed GEMM then runs with K = ld, exactly. */
int tvts_patch_gather_ld(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p, int64_t n,
                         int64_t ld, void* stream);
int tvts_cast_bf16_pad(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld, void* stream);
/* TVTS v1 front end (v1/model/video_encoder.py:78-99,178-217): Conv3d tubelets of 2 frames with an independent keep mask per tube.
 *   tubelet_gather: video [B,T,3,R,R] f32 + keep_ind [B,T/2,n] int64 -> cols [(b*T/2+tube)*n+j, ((c*2+dt)*p+u)*p+v] bf16
 *   video_assemble_tube: x0[b,0] = cls + pos[0]; x0[b,1+t*n+j] = tok[(b*nt+t)*n+j] + pos[1+keep[b,t,j]] + tem[t]  (tok includes the
 *                        Conv3d bias); _bwd: dcls/dpos/dtem ACCUMULATED, dtok written as bf16
 *   relu_bf16 / relu_bwd: the ReLU in front of txt_proj's Linear (v1/model/model_dist_TVTS.py:66-69) */
int tvts_tubelet_gather(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p, int64_t n,
                        void* stream);
int tvts_video_assemble_tube(const float* tok, const float* cls, const float* pos, const float* tem, const int64_t* keep_ind, float* x0,
                             int64_t B, int64_t nt, int64_t n, int64_t D, void* stream);
int tvts_video_assemble_tube_bwd(const float* dx0, const int64_t* keep_ind, float* dcls, float* dpos, float* dtem, void* dtok_bf16,
                                 int64_t B, int64_t nt, int64_t n, int64_t D, void* stream);
int tvts_relu_bf16(const float* x, void* y, int64_t n, void* stream);
int tvts_relu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream);
/* Input stage fused into the gather: video [B,T,3,R,R] UINT8 crops; the reference's x/255 then (x-mean[c])/std[c]
 * (v2/video_transforms/video_transform.py:24-76,627-650; mean3 / std3 are HOST pointers to 3 floats, videoaug.py:16) are applied on the
 * fly with IEEE fp32 operations in the reference's order, so cols is bit-identical to patch_gather on the normalised fp32 clip. */
int tvts_patch_gather_u8(const void* video_u8, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p, int64_t n,
                         const float* mean3, const float* std3, void* stream);
int tvts_video_assemble(const float* tok, const float* cls, const float* pos, const float* tem, const int64_t* keep_ind, float* x0,
                        int64_t B, int64_t T, int64_t n, int64_t D, void* stream);
int tvts_video_assemble_bwd(const float* dx0, const int64_t* keep_ind, float* dcls, float* dpos, float* dtem, void* dtok_bf16, int64_t B,
                            int64_t T, int64_t n, int64_t D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Text front end / pooling (v2/model/model_dist_TVTSv2_ViT_B_16.py:97-111).  tokens int32 or int64 [rows, L].
 *   text_embed: x[r,l] = table[tok[r,l]] + pos[l];  bwd: dtable scatter-add, dpos column sums (both ACCUMULATED, optional)
 *   argmax_rows: flat_idx[r] = r*L + argmax_l tok[r,l] (first maximum = EOT position, exact)
 *   gather_rows / scatter_rows: f32 rows by int64 row index (EOT pooling, sort-head transcript rows)
 *   group_mean: t [nt*B, E] clip-major -> mean over the nt transcripts (:74-76) and its backward
 */
int tvts_text_embed(const void* tokens, int64_t tok_is_i64, const float* table, const float* pos, float* x, int64_t rows, int64_t L,
                    int64_t W, void* stream);
int tvts_text_embed_bwd(const float* dx, const void* tokens, int64_t tok_is_i64, float* dtable, float* dpos, int64_t rows, int64_t L,
                        int64_t W, void* stream);
int tvts_argmax_rows(const void* tokens, int64_t tok_is_i64, int64_t* flat_idx, int64_t rows, int64_t L, void* stream);
int tvts_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t rows, int64_t D, void* stream);
int tvts_scatter_rows(const float* src, const int64_t* idx, float* dst, int64_t rows, int64_t D, int64_t accumulate, void* stream);
int tvts_group_mean(const float* t, float* out, int64_t nt, int64_t B, int64_t E, void* stream);
int tvts_group_mean_bwd(const float* dout, float* dt, void* dt_bf16, int64_t nt, int64_t B, int64_t E, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sort head glue (v2/model/sort_transformer.py:124-142).
 *   sort_concat: z[b,i<N] = vtok[b,i] + type_embed[0]; z[b,N+tr] = t[tr*B+b] + type_embed[1]
 *   sort_concat_bwd: dvtok written, dtype_embed ACCUMULATED
 *   small_linear: the E -> n_trans classifier head in fp32 (fwd; bwd: dx written, dw/db ACCUMULATED)
 */
int tvts_sort_concat(const float* vtok, const float* t, const float* type_embed, float* z, int64_t B, int64_t N, int64_t nt, int64_t E,
                     void* stream);
int tvts_sort_concat_bwd(const float* dz, float* dvtok, float* dtype_embed, int64_t B, int64_t N, int64_t nt, int64_t E, void* stream);
int tvts_small_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t R, int64_t K, int64_t O, void* stream);
int tvts_small_linear_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t R, int64_t K,
                          int64_t O, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise helpers around the GEMMs: fp32->bf16 operand cast, bias gradient (column sums, ACCUMULATED),
 * strided row add.
 */
int tvts_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
int tvts_colsum_bf16(const void* x, float* out, int64_t M, int64_t N, int64_t ld, void* stream);
int tvts_add_rows(const float* src, float* dst, int64_t rows, int64_t D, int64_t ld_src, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses (fp32).  sim_matrix = v2/model/model_dist_TVTSv2_ViT_B_16.py:119-127; NormSoftmaxLoss = v2/model/loss.py:13-25;
 * sort CE = v2/trainer/trainer.py:487-492 (weight 2, int64 labels).  gout (device scalar, optional) = upstream gradient.
 *   normalize_rows: xn = x / max(||x||, eps), norm saved
 *   sim_matrix: S[Ra,Rb] = scale * an . bn^T;  sim_matrix_bwd: gradient w.r.t. rows [row0,row0+nrows) of one side
 *               (transposed=0: the `a` side with coefficients G[row,:]; 1: the `b` side with G[:,row])
 *   nsl_fwd: loss = -mean diag log_softmax(S/T, rows) - mean diag log_softmax(S/T, cols); nsl_bwd: G = dloss/dS
 *   sort_ce: loss (optional) and dlogits (optional)
 */
int tvts_normalize_rows(const float* x, float* xn, float* norm, int64_t rows, int64_t E, float eps, void* stream);
int tvts_sim_matrix(const float* an, const float* bn, float* S, int64_t Ra, int64_t Rb, int64_t E, float scale, void* stream);
int tvts_sim_matrix_bwd(const float* G, const float* self_n, const float* other_n, const float* self_norm, float* dself, int64_t R_self,
                        int64_t R_other, int64_t E, int64_t row0, int64_t nrows, int64_t transposed, float scale, float eps, void* stream);
int tvts_nsl_fwd(const float* S, float* lse_r, float* lse_c, float* loss, int64_t Bg, float temperature, void* stream);
int tvts_nsl_bwd(const float* S, const float* lse_r, const float* lse_c, const float* gout, float* G, int64_t Bg, float temperature,
                 void* stream);
int tvts_sort_ce(const float* logits, const int64_t* labels, const float* gout, float* loss, float* dlogits, int64_t R, int64_t C,
                 float weight, void* stream);
/* All of the above in ONE launch (csrc/loss_fused.cu; a cluster of up to 8 CTAs, S and dloss/dS in distributed shared memory): from the
 * GATHERED embeddings video_all / text_all [Bg, E] (Bg <= 256, E <= 1024: tvts_contrastive_sortce_fused_supported) to loss1 =
 * NormSoftmaxLoss(sim_matrix(video_all, text_all) / temperature), its gradient w.r.t. rows [row0, row0 + nloc) of both embeddings
 * (d_video, d_text [nloc, E]: AllGather_multi's local-slice backward, v2/trainer/trainer.py:53-57), and -- when logits != NULL -- loss2 =
 * ce_weight * mean CE(logits [R, C], labels) with dlogits.  Gradients are for unit upstream gradients. */
int tvts_contrastive_sortce_fused_supported(int64_t Bg, int64_t E);
int tvts_contrastive_sortce_fused(const float* video_all, const float* text_all, int64_t Bg, int64_t E, int64_t row0, int64_t nloc,
                                  float temperature, float eps, const float* logits, const int64_t* labels, int64_t R, int64_t C,
                                  float ce_weight, float* loss1, float* loss2, float* d_video, float* d_text, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer: transformers==4.10.2 AdamW (v2/train_dist_TVTSv2_ViT_B_16.py:119-125) over one flat arena, single launch.
 * p/g/m/v are fp32 arenas with identical layout (tensors padded to whole chunks of chunk_elems); p_bf16 (optional) receives
 * the bf16 copy of the updated weights; chunk_tensor[c] = tensor index of chunk c; table[t] = {lr*sqrt(1-b2^t)/(1-b1^t),
 * lr*weight_decay, active (0 = skip: no gradient this step), 0}.  g is multiplied by grad_scale first.
 */
int tvts_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                    int64_t n_chunks, int64_t chunk_elems, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* The same update under DYNAMIC loss scaling (IEEE-half operand build), decided entirely on the device so that the step stays one
 * CUDA graph: a finite check over the gradient arena, the update (skipped as a whole when a non-finite gradient was seen; gradients
 * are divided by the current scale; bias correction evaluated from the device-side per-tensor step counters `steps`), and the scale
 * policy of torch.cuda.amp.GradScaler (x0.5 + skip on overflow, x2 after `growth_interval` finite steps, capped at max_scale).
 * table[t] = {lr, lr*weight_decay, active, correct_bias};  state = {scale, finite_steps, found_inf, skipped_steps} (fp32[4]). */
int tvts_adamw_flat_dyn(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                        int32_t* steps, float* state, int64_t n_tensors, int64_t n_chunks, int64_t chunk_elems, float beta1, float beta2,
                        float eps, float growth_interval, float max_scale, void* stream);

/* The three phases of tvts_adamw_flat_dyn as separate launches, for a data-parallel step that updates the arena bucket by bucket while
 * the NEXT bucket's gradient all-reduce is still on the wire (trainer.TrainStep, TVTS_PIPELINED_ADAMW):
 *   _check  : finite check over g[0, n_elems) -> state[2] (run on the LOCAL gradients before the all-reduce; the flag is then
 *             MAX-reduced across the ranks -- a sum of finite fp32 gradients stays finite short of 3.4e38);
 *   _apply  : the update of chunks [0, n_chunks) of the arenas passed in (pass pointers offset to the bucket's first chunk; no-op
 *             when state[2] is set);
 *   _finish : the scale policy + the per-tensor step counters (once per step, after every bucket). */
int tvts_adamw_dyn_check(const float* g, int64_t n_elems, float* state, void* stream);
int tvts_adamw_dyn_apply(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunk_tensor, const float* table,
                         const int32_t* steps, const float* state, int64_t n_chunks, int64_t chunk_elems, float beta1, float beta2, float eps,
                         void* stream);
int tvts_adamw_dyn_finish(int32_t* steps, const float* table, float* state, int64_t n_tensors, float growth_interval, float max_scale,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * The two exchange steps of the data-parallel path over NCCL (one process per GPU, one communicator per process):
 *   tvts_comm_allgather : every rank's `bytes_per_rank` bytes, concatenated in rank order -- the [B_local, 2E] embeddings of
 *                         AllGather_multi.forward (v2/trainer/trainer.py:41-57, :481-482)
 *   tvts_comm_allreduce : in-place sum (average != 0: mean) of n floats -- DistributedDataParallel's gradient averaging
 *                         (v2/base/base_trainer.py:23-25) on the flat gradient arena
 * Rank 0 creates the 128-byte id with tvts_comm_unique_id and hands it to the other ranks by any side channel (a TCP store, a file);
 * tvts_comm_init is collective and binds the communicator to the caller's current CUDA device.  Collectives are enqueued on `stream`
 * (capturable in a CUDA graph); nothing synchronises.  NCCL is resolved at run time (the libnccl.so.2 the process has loaded, else the
 * system one).  The host side uses these when TVTS_COMM=native; by default the same two collectives are torch.distributed calls.
 */
#define TVTS_COMM_ID_BYTES 128
typedef struct tvts_comm tvts_comm;
int tvts_comm_unique_id(void* id_out);
int tvts_comm_init(tvts_comm** comm, const void* id_bytes, int64_t rank, int64_t world);
int tvts_comm_allgather(tvts_comm* comm, const void* send, void* recv, int64_t bytes_per_rank, void* stream);
int tvts_comm_allreduce(tvts_comm* comm, float* buf, int64_t n, int64_t average, void* stream);
int tvts_comm_destroy(tvts_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* TVTS_B200_H */
