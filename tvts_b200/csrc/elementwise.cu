// HBM-bound glue kernels of the TVTS hot path (SURVEY.md K1/K2/K10/K12 and the autograd bookkeeping around the
// GEMMs).  All are simple streaming kernels: 128-bit vectorised coalesced global accesses, grid sized from the
// problem (>= several waves of 148 SMs at the benchmark shapes), fp32 arithmetic, exact integer indexing.
#include "common.cuh"
#include "../../include/tvts_b200.h"

namespace {

inline unsigned grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 0x7fffffffLL) g = 0x7fffffffLL;
  return (unsigned)g;
}

// ---------------------------------------------------------------- cast fp32 -> bf16
__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n4, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (i == 0)
    for (long long j = n4 * 4; j < n; ++j) dst[j] = opnd_from_float(src[j]);
}

// ---------------------------------------------------------------- column sums of a bf16 matrix (bias gradients)
// grid (ceil(N/256), row_chunks); thread = 8 consecutive columns of one row slice.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long M, long long N,
                                                     long long ld, int rows_per_cta) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long col = (long long)blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(r0 + (long long)rows_per_cta, M);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      uint4 p = *reinterpret_cast<const uint4*>(x + r * ld + col);
      float2 a = unpack_bf16x2(p.x), b = unpack_bf16x2(p.y), c = unpack_bf16x2(p.z), d = unpack_bf16x2(p.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// ---------------------------------------------------------------- patch gather (im2col of the KEPT patches only)
// video [B,T,3,R,R] fp32, keep_ind [B,n] int64 -> cols [(b*T+t)*n + j, c*p*p + u*p + v] bf16.
// Conv2d(k=s=p, no bias) of video_encoder_ViT_B_16.py:180-184 followed by the tube-mask gather :200-216 is a
// per-patch independent linear map, so masked patches are never embedded.
__global__ void patch_gather_kernel(const float* __restrict__ video, const long long* __restrict__ keep, bf16* __restrict__ cols,
                                    int B, int T, int R, int p, int n, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int pv = p / 4;                 // float4 per patch row
  const int K4 = 3 * p * pv;            // float4 per output row
  const long long row = i / K4;
  int k4 = (int)(i - row * K4);
  const int j = (int)(row % n);
  const long long bt = row / n;
  const int b = (int)(bt / T);
  const int c = k4 / (p * pv);
  k4 -= c * p * pv;
  const int u = k4 / pv, v4 = k4 - u * pv;
  const int g = R / p;
  const long long pi = keep[(long long)b * n + j];
  const int py = (int)(pi / g), px = (int)(pi % g);
  const float4 val = *reinterpret_cast<const float4*>(video + ((bt * 3 + c) * R + (py * p + u)) * (long long)R + px * p + v4 * 4);
  reinterpret_cast<uint2*>(cols)[i] = make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
}

// ---------------------------------------------------------------- video token assembly (+ its backward)
// x0[b,0] = cls + pos[0];  x0[b,1+t*n+j] = tok[(b*T+t)*n+j] + pos[1+keep[b,j]] + tem[t]   (:185-216)
__global__ void video_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                                      const float* __restrict__ tem, const long long* __restrict__ keep, float* __restrict__ x0,
                                      int B, int T, int n, int D4, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int N = 1 + T * n;
  const int c = (int)(i % D4);
  const long long rowg = i / D4;
  const int tokpos = (int)(rowg % N);
  const int b = (int)(rowg / N);
  float4 o;
  if (tokpos == 0) {
    float4 a = reinterpret_cast<const float4*>(cls)[c], q = reinterpret_cast<const float4*>(pos)[c];
    o = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  } else {
    const int t = (tokpos - 1) / n, j = (tokpos - 1) % n;
    const long long pi = keep[(long long)b * n + j];
    float4 a = reinterpret_cast<const float4*>(tok)[(((long long)b * T + t) * n + j) * D4 + c];
    float4 q = reinterpret_cast<const float4*>(pos)[(1 + pi) * D4 + c];
    float4 e = reinterpret_cast<const float4*>(tem)[(long long)t * D4 + c];
    o = make_float4(a.x + q.x + e.x, a.y + q.y + e.y, a.z + q.z + e.z, a.w + q.w + e.w);
  }
  reinterpret_cast<float4*>(x0)[i] = o;
}

// grid (T+1, B): block t<T handles frame t of sample b (thread = one float4 column, loops over the n kept slots);
// block T handles the CLS row.  dtem/dcls reduce in registers first; dpos is a scatter-add (distinct addresses per j).
__global__ void video_assemble_bwd_kernel(const float* __restrict__ dx0, const long long* __restrict__ keep, float* __restrict__ dcls,
                                          float* __restrict__ dpos, float* __restrict__ dtem, bf16* __restrict__ dtok, int T, int n,
                                          int D4) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int N = 1 + T * n;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    if (t == T) {
      float4 d = reinterpret_cast<const float4*>(dx0)[((long long)b * N) * D4 + c];
      float* pc = dcls + c * 4;
      float* pp = dpos + c * 4;
      atomicAdd(pc + 0, d.x); atomicAdd(pc + 1, d.y); atomicAdd(pc + 2, d.z); atomicAdd(pc + 3, d.w);
      atomicAdd(pp + 0, d.x); atomicAdd(pp + 1, d.y); atomicAdd(pp + 2, d.z); atomicAdd(pp + 3, d.w);
      continue;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      float4 d = reinterpret_cast<const float4*>(dx0)[((long long)b * N + 1 + (long long)t * n + j) * D4 + c];
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      const long long pi = keep[(long long)b * n + j];
      float* pp = dpos + ((1 + pi) * D4 + c) * 4;
      atomicAdd(pp + 0, d.x); atomicAdd(pp + 1, d.y); atomicAdd(pp + 2, d.z); atomicAdd(pp + 3, d.w);
      reinterpret_cast<uint2*>(dtok)[(((long long)b * T + t) * n + j) * D4 + c] =
          make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
    float* pt = dtem + ((long long)t * D4 + c) * 4;
    atomicAdd(pt + 0, acc.x); atomicAdd(pt + 1, acc.y); atomicAdd(pt + 2, acc.z); atomicAdd(pt + 3, acc.w);
  }
}

// ---------------------------------------------------------------- text embedding (+ backward) and EOT index
// x[r,l] = table[tok[r,l]] + pos[l]   (model_dist_TVTSv2_ViT_B_16.py:98-100)
template <typename TokT>
__global__ void text_embed_kernel(const TokT* __restrict__ tok, const float* __restrict__ table, const float* __restrict__ pos,
                                  float* __restrict__ x, int L, int W4, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int c = (int)(i % W4);
  const long long rl = i / W4;
  const int l = (int)(rl % L);
  const long long id = (long long)tok[rl];
  float4 a = reinterpret_cast<const float4*>(table)[id * W4 + c];
  float4 q = reinterpret_cast<const float4*>(pos)[(long long)l * W4 + c];
  reinterpret_cast<float4*>(x)[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
}
// grid (L): block l sums dpos[l] over rows in registers and scatter-adds the table rows.
template <typename TokT>
__global__ void text_embed_bwd_kernel(const float* __restrict__ dx, const TokT* __restrict__ tok, float* __restrict__ dtable,
                                      float* __restrict__ dpos, long long rows, int L, int W4) {
  const int l = blockIdx.x;
  for (int c = threadIdx.x; c < W4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
      float4 d = reinterpret_cast<const float4*>(dx)[(r * L + l) * W4 + c];
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      // positions after the EOT token receive an exactly-zero gradient (causal attention, EOT pooling); skipping them removes
      // thousands of contended atomics on the padding row (id 0) without changing the sum
      if (dtable && (d.x != 0.f || d.y != 0.f || d.z != 0.f || d.w != 0.f)) {
        const long long id = (long long)tok[r * L + l];
        float* pt = dtable + (id * W4 + c) * 4;
        atomicAdd(pt + 0, d.x); atomicAdd(pt + 1, d.y); atomicAdd(pt + 2, d.z); atomicAdd(pt + 3, d.w);
      }
    }
    if (dpos) {
      float* pp = dpos + ((long long)l * W4 + c) * 4;
      atomicAdd(pp + 0, acc.x); atomicAdd(pp + 1, acc.y); atomicAdd(pp + 2, acc.z); atomicAdd(pp + 3, acc.w);
    }
  }
}
// first index of the row maximum (torch.argmax semantics): the EOT position (:107-108).  One warp per row.
template <typename TokT>
__global__ void argmax_rows_kernel(const TokT* __restrict__ tok, long long* __restrict__ flat_idx, long long rows, int L) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  long long best = -0x7fffffffffffffffLL - 1;
  int bi = 0x7fffffff;
  for (int l = lane; l < L; l += 32) {
    long long v = (long long)tok[r * L + l];
    if (v > best) { best = v; bi = l; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) flat_idx[r] = r * L + bi;
}

// ---------------------------------------------------------------- row gather / scatter (fp32 rows by int64 index)
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst, int D4,
                                   long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const long long r = i / D4;
  const int c = (int)(i - r * D4);
  reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[idx[r] * D4 + c];
}
// dst[idx[r]] (+)= src[r]; indices are unique on this path (one EOT per text row, one slot per transcript)
__global__ void scatter_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst, int D4,
                                    long long total_vec, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const long long r = i / D4;
  const int c = (int)(i - r * D4);
  float4 v = reinterpret_cast<const float4*>(src)[i];
  float4* d = reinterpret_cast<float4*>(dst) + idx[r] * D4 + c;
  if (accumulate) {
    float4 o = *d;
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  *d = v;
}

// ---------------------------------------------------------------- transcript mean (+ backward broadcast)
// t [nt*B, E] clip-major (row tr*B+b) -> out[b] = mean_tr t[tr*B+b]     (:74-76)
__global__ void group_mean_kernel(const float* __restrict__ t, float* __restrict__ out, int nt, long long BE4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BE4) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tr = 0; tr < nt; ++tr) {
    float4 v = reinterpret_cast<const float4*>(t)[(long long)tr * BE4 + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float s = 1.0f / nt;
  reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * s, acc.y * s, acc.z * s, acc.w * s);
}
__global__ void group_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dt, bf16* __restrict__ dt_bf16, int nt,
                                      long long BE4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BE4 * nt) return;
  float4 v = reinterpret_cast<const float4*>(dout)[i % BE4];
  const float s = 1.0f / nt;
  v.x *= s; v.y *= s; v.z *= s; v.w *= s;
  if (dt) reinterpret_cast<float4*>(dt)[i] = v;
  if (dt_bf16) reinterpret_cast<uint2*>(dt_bf16)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// ---------------------------------------------------------------- sort-head input assembly (+ backward)
// z[b,i<N] = vtok[b,i] + te[0];  z[b,N+tr] = t[tr*B+b] + te[1]      (sort_transformer.py:124-129)
__global__ void sort_concat_kernel(const float* __restrict__ vtok, const float* __restrict__ t, const float* __restrict__ te,
                                   float* __restrict__ z, int B, int N, int nt, int E4, long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int S = N + nt;
  const int c = (int)(i % E4);
  const long long rs = i / E4;
  const int s = (int)(rs % S);
  const int b = (int)(rs / S);
  float4 a, e;
  if (s < N) {
    a = reinterpret_cast<const float4*>(vtok)[((long long)b * N + s) * E4 + c];
    e = reinterpret_cast<const float4*>(te)[c];
  } else {
    a = reinterpret_cast<const float4*>(t)[((long long)(s - N) * B + b) * E4 + c];
    e = reinterpret_cast<const float4*>(te)[E4 + c];
  }
  reinterpret_cast<float4*>(z)[i] = make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w);
}
// grid (chunks, B): d_vtok[b,i] = dz[b,i] (+ optional add into existing), dte[0] += sum_i dz[b,i<N], dte[1] += sum_tr dz[b,N+tr]
__global__ void sort_concat_bwd_kernel(const float* __restrict__ dz, float* __restrict__ dvtok, float* __restrict__ dte, int N, int nt,
                                       int E4, int rows_per_cta) {
  const int b = blockIdx.y;
  const int S = N + nt;
  const int s0 = blockIdx.x * rows_per_cta;
  const int s1 = min(s0 + rows_per_cta, S);
  for (int c = threadIdx.x; c < E4; c += blockDim.x) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int s = s0; s < s1; ++s) {
      float4 d = reinterpret_cast<const float4*>(dz)[((long long)b * S + s) * E4 + c];
      if (s < N) {
        reinterpret_cast<float4*>(dvtok)[((long long)b * N + s) * E4 + c] = d;
        a0.x += d.x; a0.y += d.y; a0.z += d.z; a0.w += d.w;
      } else {
        a1.x += d.x; a1.y += d.y; a1.z += d.z; a1.w += d.w;
      }
    }
    float* p0 = dte + c * 4;
    float* p1 = dte + (E4 + c) * 4;
    atomicAdd(p0 + 0, a0.x); atomicAdd(p0 + 1, a0.y); atomicAdd(p0 + 2, a0.z); atomicAdd(p0 + 3, a0.w);
    if (s1 > N) { atomicAdd(p1 + 0, a1.x); atomicAdd(p1 + 1, a1.y); atomicAdd(p1 + 2, a1.z); atomicAdd(p1 + 3, a1.w); }
  }
}

// ---------------------------------------------------------------- strided row add: dst[r*ldd + c] += src[r*lds + c]
__global__ void add_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long lds4, long long ldd4, int D4,
                                long long total_vec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const long long r = i / D4;
  const int c = (int)(i - r * D4);
  float4 v = reinterpret_cast<const float4*>(src)[r * lds4 + c];
  float4* d = reinterpret_cast<float4*>(dst) + r * ldd4 + c;
  float4 o = *d;
  *d = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
}

// ---------------------------------------------------------------- tiny fp32 linear (sort head E -> n_trans) + backward
// y[r,o] = b[o] + sum_k x[r,k] w[o,k]; one warp per (row, out) pair; O is tiny (n_trans = 4).
__global__ void small_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                        float* __restrict__ y, long long R, int K, int O) {
  const long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wi >= R * O) return;
  const long long r = wi / O;
  const int o = (int)(wi - r * O);
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += x[r * K + k] * w[(long long)o * K + k];
  s = warp_sum(s);
  if (lane == 0) y[wi] = s + (bias ? bias[o] : 0.f);
}
// dx[r,k] = sum_o dy[r,o] w[o,k];  dw[o,k] += sum_r dy[r,o] x[r,k];  db[o] += sum_r dy[r,o].   grid (ceil(K/256)), thread = column k
__global__ void small_linear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                        float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long R, int K, int O) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  for (int o = 0; o < O; ++o) {
    float acc = 0.f, accb = 0.f;
    for (long long r = blockIdx.y; r < R; r += gridDim.y) {
      const float g = dy[r * O + o];
      acc += g * x[r * K + k];
      accb += g;
    }
    atomicAdd(dw + (long long)o * K + k, acc);
    if (k == 0 && db) atomicAdd(db + o, accb);
  }
  for (long long r = blockIdx.y; r < R; r += gridDim.y) {
    float s = 0.f;
    for (int o = 0; o < O; ++o) s += dy[r * O + o] * w[(long long)o * K + k];
    dx[r * K + k] = s;
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tvts_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return TVTS_OK;
  TVTS_REQUIRE(src && dst && n > 0, "cast_bf16: bad arguments");
  TVTS_REQUIRE((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 8 == 0, "cast_bf16: alignment");
  const long long n4 = n / 4;
  cast_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, ST(stream)>>>(src, (bf16*)dst, n4, n);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_colsum_bf16(const void* x, float* out, int64_t M, int64_t N, int64_t ld, void* stream) {
  if (M == 0 || N == 0) return TVTS_OK;
  TVTS_REQUIRE(x && out && N % 8 == 0 && ld % 8 == 0, "colsum_bf16: N and ld must be multiples of 8");
  const int gx = (int)((N + 255) / 256);
  long long chunks = (4LL * tvts_num_sms() + gx - 1) / gx;
  if (chunks > (M + 63) / 64) chunks = (M + 63) / 64;
  if (chunks < 1) chunks = 1;
  const int rows_per = (int)((M + chunks - 1) / chunks);
  dim3 grid(gx, (unsigned)((M + rows_per - 1) / rows_per));
  colsum_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)x, out, M, N, ld, rows_per);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_patch_gather(const float* video, const int64_t* keep_ind, void* cols, int64_t B, int64_t T, int64_t R, int64_t p,
                                 int64_t n, void* stream) {
  TVTS_REQUIRE(video && keep_ind && cols, "patch_gather: null pointer");
  TVTS_REQUIRE(p % 4 == 0 && R % p == 0, "patch_gather: patch=%lld must be a multiple of 4 and divide the resolution", (long long)p);
  const long long total = B * T * n * 3 * p * (p / 4);
  if (total == 0) return TVTS_OK;
  patch_gather_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(video, (const long long*)keep_ind, (bf16*)cols, (int)B, (int)T,
                                                                    (int)R, (int)p, (int)n, total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_video_assemble(const float* tok, const float* cls, const float* pos, const float* tem, const int64_t* keep_ind,
                                   float* x0, int64_t B, int64_t T, int64_t n, int64_t D, void* stream) {
  TVTS_REQUIRE(tok && cls && pos && tem && keep_ind && x0 && D % 4 == 0, "video_assemble: bad arguments");
  const long long total = B * (1 + T * n) * (D / 4);
  if (total == 0) return TVTS_OK;
  video_assemble_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(tok, cls, pos, tem, (const long long*)keep_ind, x0, (int)B,
                                                                      (int)T, (int)n, (int)(D / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_video_assemble_bwd(const float* dx0, const int64_t* keep_ind, float* dcls, float* dpos, float* dtem, void* dtok_bf16,
                                       int64_t B, int64_t T, int64_t n, int64_t D, void* stream) {
  TVTS_REQUIRE(dx0 && keep_ind && dcls && dpos && dtem && dtok_bf16 && D % 4 == 0, "video_assemble_bwd: bad arguments");
  if (B == 0) return TVTS_OK;
  dim3 grid((unsigned)(T + 1), (unsigned)B);
  video_assemble_bwd_kernel<<<grid, 192, 0, ST(stream)>>>(dx0, (const long long*)keep_ind, dcls, dpos, dtem, (bf16*)dtok_bf16, (int)T,
                                                          (int)n, (int)(D / 4));
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_text_embed(const void* tokens, int64_t tok_is_i64, const float* table, const float* pos, float* x, int64_t rows,
                               int64_t L, int64_t W, void* stream) {
  TVTS_REQUIRE(tokens && table && pos && x && W % 4 == 0, "text_embed: bad arguments");
  const long long total = rows * L * (W / 4);
  if (total == 0) return TVTS_OK;
  if (tok_is_i64)
    text_embed_kernel<long long><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const long long*)tokens, table, pos, x, (int)L,
                                                                               (int)(W / 4), total);
  else
    text_embed_kernel<int><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const int*)tokens, table, pos, x, (int)L, (int)(W / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_text_embed_bwd(const float* dx, const void* tokens, int64_t tok_is_i64, float* dtable, float* dpos, int64_t rows,
                                   int64_t L, int64_t W, void* stream) {
  TVTS_REQUIRE(dx && tokens && W % 4 == 0, "text_embed_bwd: bad arguments");
  if (rows == 0) return TVTS_OK;
  long long gy = (4LL * tvts_num_sms() + L - 1) / L;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  dim3 grid((unsigned)L, (unsigned)gy);
  if (tok_is_i64)
    text_embed_bwd_kernel<long long><<<grid, 128, 0, ST(stream)>>>(dx, (const long long*)tokens, dtable, dpos, rows, (int)L, (int)(W / 4));
  else
    text_embed_bwd_kernel<int><<<grid, 128, 0, ST(stream)>>>(dx, (const int*)tokens, dtable, dpos, rows, (int)L, (int)(W / 4));
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_argmax_rows(const void* tokens, int64_t tok_is_i64, int64_t* flat_idx, int64_t rows, int64_t L, void* stream) {
  TVTS_REQUIRE(tokens && flat_idx && L > 0, "argmax_rows: bad arguments");
  if (rows == 0) return TVTS_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (tok_is_i64) argmax_rows_kernel<long long><<<grid, 256, 0, ST(stream)>>>((const long long*)tokens, (long long*)flat_idx, rows, (int)L);
  else argmax_rows_kernel<int><<<grid, 256, 0, ST(stream)>>>((const int*)tokens, (long long*)flat_idx, rows, (int)L);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t rows, int64_t D, void* stream) {
  TVTS_REQUIRE(src && idx && dst && D % 4 == 0, "gather_rows: bad arguments");
  const long long total = rows * (D / 4);
  if (total == 0) return TVTS_OK;
  gather_rows_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(src, (const long long*)idx, dst, (int)(D / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_scatter_rows(const float* src, const int64_t* idx, float* dst, int64_t rows, int64_t D, int64_t accumulate,
                                 void* stream) {
  TVTS_REQUIRE(src && idx && dst && D % 4 == 0, "scatter_rows: bad arguments");
  const long long total = rows * (D / 4);
  if (total == 0) return TVTS_OK;
  scatter_rows_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(src, (const long long*)idx, dst, (int)(D / 4), total, (int)accumulate);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_group_mean(const float* t, float* out, int64_t nt, int64_t B, int64_t E, void* stream) {
  TVTS_REQUIRE(t && out && nt > 0 && E % 4 == 0, "group_mean: bad arguments");
  const long long be4 = B * (E / 4);
  if (be4 == 0) return TVTS_OK;
  group_mean_kernel<<<grid_for(be4, 256), 256, 0, ST(stream)>>>(t, out, (int)nt, be4);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_group_mean_bwd(const float* dout, float* dt, void* dt_bf16, int64_t nt, int64_t B, int64_t E, void* stream) {
  TVTS_REQUIRE(dout && nt > 0 && E % 4 == 0, "group_mean_bwd: bad arguments");
  const long long be4 = B * (E / 4);
  if (be4 == 0) return TVTS_OK;
  group_mean_bwd_kernel<<<grid_for(be4 * nt, 256), 256, 0, ST(stream)>>>(dout, dt, (bf16*)dt_bf16, (int)nt, be4);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_sort_concat(const float* vtok, const float* t, const float* type_embed, float* z, int64_t B, int64_t N, int64_t nt,
                                int64_t E, void* stream) {
  TVTS_REQUIRE(vtok && t && type_embed && z && E % 4 == 0, "sort_concat: bad arguments");
  const long long total = B * (N + nt) * (E / 4);
  if (total == 0) return TVTS_OK;
  sort_concat_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(vtok, t, type_embed, z, (int)B, (int)N, (int)nt, (int)(E / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_sort_concat_bwd(const float* dz, float* dvtok, float* dtype_embed, int64_t B, int64_t N, int64_t nt, int64_t E,
                                    void* stream) {
  TVTS_REQUIRE(dz && dvtok && dtype_embed && E % 4 == 0, "sort_concat_bwd: bad arguments");
  if (B == 0) return TVTS_OK;
  const int S = (int)(N + nt);
  const int rows_per = 32;
  dim3 grid((unsigned)((S + rows_per - 1) / rows_per), (unsigned)B);
  sort_concat_bwd_kernel<<<grid, 128, 0, ST(stream)>>>(dz, dvtok, dtype_embed, (int)N, (int)nt, (int)(E / 4), rows_per);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_add_rows(const float* src, float* dst, int64_t rows, int64_t D, int64_t ld_src, int64_t ld_dst, void* stream) {
  TVTS_REQUIRE(src && dst && D % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, "add_rows: bad arguments");
  const long long total = rows * (D / 4);
  if (total == 0) return TVTS_OK;
  add_rows_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(src, dst, ld_src / 4, ld_dst / 4, (int)(D / 4), total);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_small_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t R, int64_t K, int64_t O,
                                     void* stream) {
  TVTS_REQUIRE(x && w && y, "small_linear_fwd: bad arguments");
  if (R * O == 0) return TVTS_OK;
  small_linear_fwd_kernel<<<(unsigned)((R * O + 7) / 8), 256, 0, ST(stream)>>>(x, w, bias, y, R, (int)K, (int)O);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}

extern "C" int tvts_small_linear_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t R,
                                     int64_t K, int64_t O, void* stream) {
  TVTS_REQUIRE(dy && x && w && dx && dw, "small_linear_bwd: bad arguments");
  if (R == 0) return TVTS_OK;
  // rows are strided over gridDim.y: every (r, k) of dx is written by exactly one block, dw / db partials meet in atomics
  dim3 grid((unsigned)((K + 255) / 256), (unsigned)(R < 32 ? R : 32));
  small_linear_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(dy, x, w, dx, dw, db, R, (int)K, (int)O);
  TVTS_LAUNCH_CHECK();
  return TVTS_OK;
}
