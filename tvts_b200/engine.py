"""Host-side orchestration of the TVTSv2 hot path on top of the C-ABI kernels (tvts_b200/_lib.py).

Each tower is ONE autograd node with a hand-written backward (no autograd tape inside): the forward enqueues the
kernels and keeps the activations the backward needs; the backward enqueues dgrad / wgrad / reduction kernels and
returns the parameter gradients in reference-parameter order, so nn.Module / DDP / optimizers see ordinary
nn.Parameters with ordinary .grad tensors.

Numerics: fp32 master weights, fp32 residual stream / LayerNorm statistics / softmax / losses; bf16 GEMM operands with
fp32 accumulation (tcgen05, TMEM); bf16 qkv / attention-output / MLP-hidden activations and bf16 branch gradients (the fp32
residual-gradient stream is kept in fp32).

Reference semantics restated here (paths relative to /root/reference):
  video tower  v2/model/video_encoder_ViT_B_16.py:176-235 (+ block :113-124, VarAttention :38-76)
  text tower   v2/model/model_dist_TVTSv2_ViT_B_16.py:97-111, v2/CLIP/clip/model.py:171-203
  sort head    v2/model/sort_transformer.py:124-142
  losses       v2/model/loss.py:13-25, model_dist_TVTSv2_ViT_B_16.py:119-127, v2/trainer/trainer.py:481-494
"""
import weakref

import torch

from . import _lib as L

BF16 = L.OPERAND_DTYPE          # the 16-bit operand format of the loaded library: bfloat16 (default) or float16 (TVTS_OPERAND=fp16)
F32 = torch.float32

MODE_FULL, MODE_SPACE, MODE_TIME = 0, 1, 2


# --------------------------------------------------------------------------------------------------
# bf16 operand cache for the fp32 master weights (re-cast only when the parameter changed)
# --------------------------------------------------------------------------------------------------
class _WeightCache:
    """bf16 operand copies of fp32 parameters that are not in a FlatState arena (frozen weights, optimizer-less runs).
    An entry is valid only for the very same Parameter object (weak reference: `id()` values are recycled once a model is
    garbage-collected), the same storage and the same version counter."""

    def __init__(self):
        self._c = {}

    def get(self, p):
        fs = _flat()
        if fs is not None and fs.has(p):
            return fs.bf16_view(p)          # arena copy, kept current by the fused optimizer kernel
        key = id(p)
        ent = self._c.get(key)
        ver = p._version
        if ent is not None and ent[3]() is p and ent[0] == ver and ent[1] == p.data_ptr() and ent[2].shape == p.shape \
                and ent[2].device == p.device:
            return ent[2]
        src = p.detach()
        if not src.is_contiguous():
            src = src.contiguous()
        reuse = ent is not None and ent[3]() is p and ent[2].shape == p.shape and ent[2].device == p.device
        dst = ent[2] if reuse else torch.empty(src.shape, dtype=BF16, device=p.device)
        L.call("cast_bf16", src, dst, src.numel())
        if len(self._c) > 4096:             # drop entries of dead parameters
            self._c = {k: v for k, v in self._c.items() if v[3]() is not None}
        self._c[key] = (ver, p.data_ptr(), dst, weakref.ref(p))
        return dst

    def refresh(self):
        """Re-cast, into the SAME buffers, the copies of parameters that were modified since they were cast (checkpoint resume,
        manual edits): a captured CUDA graph reads those buffers directly and never comes back through get()."""
        n = 0
        for key, (ver, ptr, dst, ref) in list(self._c.items()):
            p = ref()
            if p is None or p._version == ver or p.data_ptr() != ptr or p.shape != dst.shape:
                continue
            L.call("cast_bf16", p.detach().contiguous(), dst, p.numel())
            self._c[key] = (p._version, ptr, dst, ref)
            n += 1
        return n

    def clear(self):
        self._c.clear()


WEIGHTS = _WeightCache()


def _flat():
    from . import optim
    return optim.current()


# --------------------------------------------------------------------------------------------------
# thin op wrappers (allocate outputs, call the C ABI)
# --------------------------------------------------------------------------------------------------
def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _zeros(shape, dtype, like):
    return torch.zeros(shape, dtype=dtype, device=like.device)


def ln_fwd(x, w, b, eps, out_dtype=BF16):
    M, D = x.shape
    y = _empty((M, D), out_dtype, x)
    mean = _empty((M,), F32, x)
    rstd = _empty((M,), F32, x)
    L.call("layernorm_fwd", x, w, b, y, int(out_dtype == BF16), mean, rstd, M, D, float(eps))
    return y, mean, rstd


def ln_bwd(dy, x, mean, rstd, w, res1=None, res2=None, want_f32=True, want_bf16=True, dw=None, db=None, dxsum=None):
    """dxsum (optional, ACCUMULATED): column sums of dx = the bias gradient of the Linear whose output gradient dx is."""
    M, D = x.shape
    dx = _empty((M, D), F32, x) if want_f32 else None
    dxb = _empty((M, D), BF16, x) if want_bf16 else None
    if dxsum is None:
        L.call("layernorm_bwd", dy, int(dy.dtype == BF16), x, mean, rstd, w, res1, res2, dx, dxb, dw, db, M, D)
    else:
        L.call("layernorm_bwd_colsum", dy, int(dy.dtype == BF16), x, mean, rstd, w, res1, res2, dx, dxb, dw, db, dxsum, M, D)
    return dx, dxb


def lin_fwd(x, w, bias, out_dtype, act=None, residual=None, want_pre=False):
    """y = act(x @ w^T + bias) (+ residual).  x [M,K] bf16, w [N,K] bf16."""
    M, K = x.shape
    N = w.shape[0]
    out = _empty((M, N), out_dtype, x)
    pre = _empty((M, N), BF16, x) if want_pre else None
    L.gemm(x, w, out, M=M, N=N, K=K, lda=K, ldb=K, bias=bias, residual=residual, act=act, out_pre=pre)
    return (out, pre) if want_pre else out


def lin_dgrad(dy, w, out_dtype, dact=None, aux=None):
    """dx[M,K] = dy[M,N] @ w[N,K] (optionally * act'(aux)).  w is read MN-major (no transpose copy)."""
    M, N = dy.shape
    K = w.shape[1]
    out = _empty((M, K), out_dtype, dy)
    L.gemm(dy, w, out, M=M, N=K, K=N, lda=N, ldb=K, b_mn=True, dact=dact, aux=aux, ldaux=K)
    return out


def lin_wgrad(dy, x, dw):
    """dw[N,K] += dy[M,N]^T @ x[M,K]; both operands are read MN-major straight from the activations (split-K atomics)."""
    M, N = dy.shape
    K = x.shape[1]
    L.gemm(dy, x, dw, M=N, N=K, K=M, lda=N, ldb=K, a_mn=True, b_mn=True, accumulate=True)
    return dw


def mat_fwd(x, p):
    """y = x @ p for a [K,N] parameter (video `proj`, `text_projection`): p is the MN-major B operand."""
    M, K = x.shape
    N = p.shape[1]
    out = _empty((M, N), F32, x)
    L.gemm(x, p, out, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True)
    return out


def mat_dgrad(dy, p, out_dtype):
    """dx[M,K] = dy[M,N] @ p[K,N]^T: p is a K-major B operand."""
    M, N = dy.shape
    K = p.shape[0]
    out = _empty((M, K), out_dtype, dy)
    L.gemm(dy, p, out, M=M, N=K, K=N, lda=N, ldb=N)
    return out


def mat_wgrad(x, dy, dp):
    """dp[K,N] += x[M,K]^T @ dy[M,N]"""
    M, K = x.shape
    N = dy.shape[1]
    L.gemm(x, dy, dp, M=K, N=N, K=M, lda=K, ldb=N, a_mn=True, b_mn=True, accumulate=True)
    return dp


def colsum(dy, out):
    M, N = dy.shape
    L.call("colsum_bf16", dy, out, M, N, N)
    return out


def cast_bf16(x):
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    L.call("cast_bf16", x, y, x.numel())
    return y


def attn_fwd(qkv, B, N, H, mode, T=0, n=0, causal=False):
    d = qkv.shape[-1] // (3 * H)
    out = _empty((B * N, H * d), BF16, qkv)
    lse = _empty((B, H, N), F32, qkv)
    L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, int(causal), float(d ** -0.5))
    return out, lse


def attn_bwd(qkv, out, dout, lse, B, N, H, mode, T=0, n=0, causal=False, dbias=None):
    """dbias (optional, ACCUMULATED): the qkv Linear's bias gradient = column sums of dqkv, produced by the attention backward itself."""
    d = qkv.shape[-1] // (3 * H)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    if dbias is None:
        L.call("attn_bwd", qkv, out, dout, lse, delta, dqkv, B, N, H, d, mode, T, n, int(causal), float(d ** -0.5))
    else:
        L.call("attn_bwd_bias", qkv, out, dout, lse, delta, dqkv, dbias, B, N, H, d, mode, T, n, int(causal), float(d ** -0.5))
    return dqkv


def gather_rows_any(src, idx, rows):
    """dst[r] = src[idx[r]] for fp32 or bf16 row-major matrices (bf16 rows are moved as fp32 words)."""
    D = src.shape[1]
    dst = torch.empty((rows, D), dtype=src.dtype, device=src.device)
    if src.dtype == BF16:
        L.call("gather_rows", src.view(F32), idx, dst.view(F32), rows, D // 2)
    else:
        L.call("gather_rows", src, idx, dst, rows, D)
    return dst


def scatter_rows_into_zeros(src, idx, total_rows):
    """zeros[total_rows, D] with dst[idx[r]] = src[r] (fp32 or bf16)."""
    rows, D = src.shape
    dst = torch.zeros((total_rows, D), dtype=src.dtype, device=src.device)
    if src.dtype == BF16:
        L.call("scatter_rows", src.view(F32), idx, dst.view(F32), rows, D // 2, 0)
    else:
        L.call("scatter_rows", src, idx, dst, rows, D, 0)
    return dst


# --------------------------------------------------------------------------------------------------
# parameter access helpers
# --------------------------------------------------------------------------------------------------
class ParamView:
    """name -> tensor lookup over the flat tensor list an autograd.Function receives; collects gradients."""

    def __init__(self, names, tensors, needs):
        self.names = names
        self.t = dict(zip(names, tensors))
        self.needs = dict(zip(names, needs))
        self.grads = {}
        self.direct = set()

    def __getitem__(self, k):
        return self.t[k]

    def bf(self, k):
        return WEIGHTS.get(self.t[k])

    def need(self, k):
        return self.needs.get(k, False)

    def gbuf(self, k):
        """fp32 zero-initialised gradient accumulator for parameter k (created on first use)."""
        g = self.grads.get(k)
        if g is None:
            fs = _flat()
            if fs is not None and fs.has(self.t[k]):
                g = fs.grad_view(self.t[k])     # zeroed by FlatState.zero_grad() at the start of the step
                self.direct.add(k)
            else:
                g = torch.zeros_like(self.t[k], dtype=F32, memory_format=torch.contiguous_format)
            self.grads[k] = g
        return g

    def grad_tuple(self):
        """Gradients in parameter order for autograd; arena-backed ones are attached to .grad directly (no copy through
        AccumulateGrad) and reported as None."""
        out = []
        for k in self.names:
            g = self.grads.get(k) if self.needs.get(k, False) else None
            if g is not None and k in self.direct:
                p = self.t[k]
                if p.grad is None:
                    p.grad = g
                elif p.grad.data_ptr() != g.data_ptr():
                    p.grad.add_(g)
                g = None
            out.append(g)
        return tuple(out)


def _linear_bwd(P, wname, bname, dy_bf, x_bf):
    """weight/bias gradients of y = x W^T + b (given dy bf16 and the saved bf16 input)."""
    if P.need(wname):
        W = P[wname]
        lin_wgrad(dy_bf, x_bf, P.gbuf(wname).view(W.shape[0], -1))
    if bname is not None and P.need(bname):
        colsum(dy_bf, P.gbuf(bname))


# --------------------------------------------------------------------------------------------------
# divided space-time block (video)
# --------------------------------------------------------------------------------------------------
def st_block_fwd(P, p, x, B, N, T, n, H, act, eps):
    a3, mu3, rs3 = ln_fwd(x, P[p + "ln_3.weight"], P[p + "ln_3.bias"], eps)
    qkv_t = lin_fwd(a3, P.bf(p + "timeattn.qkv.weight"), P[p + "timeattn.qkv.bias"], BF16)
    o_t, lse_t = attn_fwd(qkv_t, B, N, H, MODE_TIME, T, n)
    tr = lin_fwd(o_t, P.bf(p + "timeattn.proj.weight"), P[p + "timeattn.proj.bias"], F32, residual=x)
    a1, mu1, rs1 = ln_fwd(tr, P[p + "ln_1.weight"], P[p + "ln_1.bias"], eps)
    qkv_s = lin_fwd(a1, P.bf(p + "attn.qkv.weight"), P[p + "attn.qkv.bias"], BF16)
    o_s, lse_s = attn_fwd(qkv_s, B, N, H, MODE_SPACE, T, n)
    sr = lin_fwd(o_s, P.bf(p + "attn.proj.weight"), P[p + "attn.proj.bias"], F32, residual=x)   # residual is x (:121)
    a2, mu2, rs2 = ln_fwd(sr, P[p + "ln_2.weight"], P[p + "ln_2.bias"], eps)
    g, h = lin_fwd(a2, P.bf(p + "mlp.c_fc.weight"), P[p + "mlp.c_fc.bias"], BF16, act=act, want_pre=True)
    out = lin_fwd(g, P.bf(p + "mlp.c_proj.weight"), P[p + "mlp.c_proj.bias"], F32, residual=sr)
    saved = (x, a3, mu3, rs3, qkv_t, o_t, lse_t, tr, a1, mu1, rs1, qkv_s, o_s, lse_s, sr, a2, mu2, rs2, h, g)
    return out, saved


def _ln_grads(P, p, name):
    need = P.need(p + name + ".weight") or P.need(p + name + ".bias")
    if not need:
        return None, None
    return P.gbuf(p + name + ".weight"), P.gbuf(p + name + ".bias")


def _bias_buf(P, name):
    return P.gbuf(name) if (name is not None and P.need(name)) else None


def st_block_bwd(P, p, saved, d_out, d_out_bf, B, N, T, n, H, act, prev_cproj_bias=None):
    """Bias gradients of the three projections that feed the residual stream are column sums of a gradient that a LayerNorm
    backward produces anyway, so they are accumulated there (ln_bwd dxsum): attn.proj <- ln_2, timeattn.proj <- ln_1, and the
    PREVIOUS block's mlp.c_proj <- this block's ln_3 (this block's own c_proj bias was filled by the producer of d_out)."""
    (x, a3, mu3, rs3, qkv_t, o_t, lse_t, tr, a1, mu1, rs1, qkv_s, o_s, lse_s, sr, a2, mu2, rs2, h, g) = saved
    # MLP
    _linear_bwd(P, p + "mlp.c_proj.weight", None, d_out_bf, g)
    dh = lin_dgrad(d_out_bf, P.bf(p + "mlp.c_proj.weight"), BF16, dact=act, aux=h)
    _linear_bwd(P, p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", dh, a2)
    da2 = lin_dgrad(dh, P.bf(p + "mlp.c_fc.weight"), BF16)   # LN backward reads its dy as bf16: half the traffic of this stream
    dw, db = _ln_grads(P, p, "ln_2")
    d_sr, d_sr_bf = ln_bwd(da2, sr, mu2, rs2, P[p + "ln_2.weight"], res1=d_out, dw=dw, db=db, dxsum=_bias_buf(P, p + "attn.proj.bias"))
    # space attention
    _linear_bwd(P, p + "attn.proj.weight", None, d_sr_bf, o_s)
    do_s = lin_dgrad(d_sr_bf, P.bf(p + "attn.proj.weight"), BF16)
    dqkv_s = attn_bwd(qkv_s, o_s, do_s, lse_s, B, N, H, MODE_SPACE, T, n, dbias=_bias_buf(P, p + "attn.qkv.bias"))
    _linear_bwd(P, p + "attn.qkv.weight", None, dqkv_s, a1)
    da1 = lin_dgrad(dqkv_s, P.bf(p + "attn.qkv.weight"), BF16)
    dw, db = _ln_grads(P, p, "ln_1")
    d_tr, d_tr_bf = ln_bwd(da1, tr, mu1, rs1, P[p + "ln_1.weight"], dw=dw, db=db,
                           dxsum=_bias_buf(P, p + "timeattn.proj.bias"))     # tr only feeds ln_1
    # time attention
    _linear_bwd(P, p + "timeattn.proj.weight", None, d_tr_bf, o_t)
    do_t = lin_dgrad(d_tr_bf, P.bf(p + "timeattn.proj.weight"), BF16)
    dqkv_t = attn_bwd(qkv_t, o_t, do_t, lse_t, B, N, H, MODE_TIME, T, n, dbias=_bias_buf(P, p + "timeattn.qkv.bias"))
    _linear_bwd(P, p + "timeattn.qkv.weight", None, dqkv_t, a3)
    da3 = lin_dgrad(dqkv_t, P.bf(p + "timeattn.qkv.weight"), BF16)
    dw, db = _ln_grads(P, p, "ln_3")
    d_x, d_x_bf = ln_bwd(da3, x, mu3, rs3, P[p + "ln_3.weight"], res1=d_sr, res2=d_tr, dw=dw, db=db, dxsum=_bias_buf(P, prev_cproj_bias))
    return d_x, d_x_bf


# --------------------------------------------------------------------------------------------------
# generic pre-LN block (CLIP text block: causal + QuickGELU; sort-head block: full + GELU)
# --------------------------------------------------------------------------------------------------
class BlockNames:
    def __init__(self, ln1, qkv_w, qkv_b, out_w, out_b, ln2, fc_w, fc_b, proj_w, proj_b):
        self.ln1, self.qkv_w, self.qkv_b, self.out_w, self.out_b = ln1, qkv_w, qkv_b, out_w, out_b
        self.ln2, self.fc_w, self.fc_b, self.proj_w, self.proj_b = ln2, fc_w, fc_b, proj_w, proj_b


def clip_block_names(p):
    return BlockNames(p + "ln_1", p + "attn.in_proj_weight", p + "attn.in_proj_bias", p + "attn.out_proj.weight",
                      p + "attn.out_proj.bias", p + "ln_2", p + "mlp.c_fc.weight", p + "mlp.c_fc.bias",
                      p + "mlp.c_proj.weight", p + "mlp.c_proj.bias")


def sort_block_names(p):
    return BlockNames(p + "norm1", p + "attn.qkv.weight", p + "attn.qkv.bias", p + "attn.proj.weight", p + "attn.proj.bias",
                      p + "norm2", p + "mlp.fc1.weight", p + "mlp.fc1.bias", p + "mlp.fc2.weight", p + "mlp.fc2.bias")


def block_fwd(P, nm, x, B, S, H, act, eps, causal):
    a1, mu1, rs1 = ln_fwd(x, P[nm.ln1 + ".weight"], P[nm.ln1 + ".bias"], eps)
    qkv = lin_fwd(a1, P.bf(nm.qkv_w), P[nm.qkv_b], BF16)
    o, lse = attn_fwd(qkv, B, S, H, MODE_FULL, causal=causal)
    x1 = lin_fwd(o, P.bf(nm.out_w), P[nm.out_b], F32, residual=x)
    a2, mu2, rs2 = ln_fwd(x1, P[nm.ln2 + ".weight"], P[nm.ln2 + ".bias"], eps)
    g, h = lin_fwd(a2, P.bf(nm.fc_w), P[nm.fc_b], BF16, act=act, want_pre=True)
    out = lin_fwd(g, P.bf(nm.proj_w), P[nm.proj_b], F32, residual=x1)
    return out, (x, a1, mu1, rs1, qkv, o, lse, x1, a2, mu2, rs2, h, g)


def block_bwd(P, nm, saved, d_out, d_out_bf, B, S, H, act, causal):
    (x, a1, mu1, rs1, qkv, o, lse, x1, a2, mu2, rs2, h, g) = saved
    _linear_bwd(P, nm.proj_w, nm.proj_b, d_out_bf, g)
    dh = lin_dgrad(d_out_bf, P.bf(nm.proj_w), BF16, dact=act, aux=h)
    _linear_bwd(P, nm.fc_w, nm.fc_b, dh, a2)
    da2 = lin_dgrad(dh, P.bf(nm.fc_w), BF16)   # LN backward reads its dy as bf16: half the traffic of this stream
    need2 = P.need(nm.ln2 + ".weight") or P.need(nm.ln2 + ".bias")
    d_x1, d_x1_bf = ln_bwd(da2, x1, mu2, rs2, P[nm.ln2 + ".weight"], res1=d_out,
                           dw=P.gbuf(nm.ln2 + ".weight") if need2 else None, db=P.gbuf(nm.ln2 + ".bias") if need2 else None,
                           dxsum=_bias_buf(P, nm.out_b))           # out-projection bias gradient = column sums of d_x1
    _linear_bwd(P, nm.out_w, None, d_x1_bf, o)
    do = lin_dgrad(d_x1_bf, P.bf(nm.out_w), BF16)
    dqkv = attn_bwd(qkv, o, do, lse, B, S, H, MODE_FULL, causal=causal, dbias=_bias_buf(P, nm.qkv_b))
    _linear_bwd(P, nm.qkv_w, None, dqkv, a1)
    da1 = lin_dgrad(dqkv, P.bf(nm.qkv_w), BF16)   # LN backward reads its dy as bf16: half the traffic of this stream
    need1 = P.need(nm.ln1 + ".weight") or P.need(nm.ln1 + ".bias")
    d_x, d_x_bf = ln_bwd(da1, x, mu1, rs1, P[nm.ln1 + ".weight"], res1=d_x1,
                         dw=P.gbuf(nm.ln1 + ".weight") if need1 else None, db=P.gbuf(nm.ln1 + ".bias") if need1 else None)
    return d_x, d_x_bf


# --------------------------------------------------------------------------------------------------
# LAST block of the sort head: only its n_trans transcript rows are consumed (norm + head, sort_transformer.py:134-142), so the
# attention queries, the projection and the MLP are evaluated for those rows alone; K / V (hence LN1 + the qkv GEMM) still cover
# every token.  Identical results, ~S / n_trans times less attention / MLP work in this block (forward and backward).
# --------------------------------------------------------------------------------------------------
def window_block_fwd(P, nm, z, idx, B, S, q0, qn, H, act, eps):
    E_ = z.shape[1]
    d = E_ // H
    a1, mu1, rs1 = ln_fwd(z, P[nm.ln1 + ".weight"], P[nm.ln1 + ".bias"], eps)
    qkv = lin_fwd(a1, P.bf(nm.qkv_w), P[nm.qkv_b], BF16)
    o = _zeros((B * S, E_), BF16, z)
    lse = _zeros((B, H, S), F32, z)
    L.call("attn_window_fwd", qkv, o, lse, B, S, H, d, q0, qn, float(d ** -0.5))
    o_t = gather_rows_any(o, idx, B * qn)
    z_t = gather_rows_any(z, idx, B * qn)
    x1 = lin_fwd(o_t, P.bf(nm.out_w), P[nm.out_b], F32, residual=z_t)
    a2, mu2, rs2 = ln_fwd(x1, P[nm.ln2 + ".weight"], P[nm.ln2 + ".bias"], eps)
    g, h = lin_fwd(a2, P.bf(nm.fc_w), P[nm.fc_b], BF16, act=act, want_pre=True)
    out = lin_fwd(g, P.bf(nm.proj_w), P[nm.proj_b], F32, residual=x1)
    return out, (z, a1, mu1, rs1, qkv, o, lse, o_t, x1, a2, mu2, rs2, h, g)


def window_block_bwd(P, nm, saved, d_out, idx, B, S, q0, qn, H, act):
    (z, a1, mu1, rs1, qkv, o, lse, o_t, x1, a2, mu2, rs2, h, g) = saved
    E_ = z.shape[1]
    d = E_ // H
    d_out_bf = cast_bf16(d_out)
    _linear_bwd(P, nm.proj_w, nm.proj_b, d_out_bf, g)
    dh = lin_dgrad(d_out_bf, P.bf(nm.proj_w), BF16, dact=act, aux=h)
    _linear_bwd(P, nm.fc_w, nm.fc_b, dh, a2)
    da2 = lin_dgrad(dh, P.bf(nm.fc_w), BF16)
    need2 = P.need(nm.ln2 + ".weight") or P.need(nm.ln2 + ".bias")
    d_x1, d_x1_bf = ln_bwd(da2, x1, mu2, rs2, P[nm.ln2 + ".weight"], res1=d_out,
                           dw=P.gbuf(nm.ln2 + ".weight") if need2 else None, db=P.gbuf(nm.ln2 + ".bias") if need2 else None)
    _linear_bwd(P, nm.out_w, nm.out_b, d_x1_bf, o_t)
    do_t = lin_dgrad(d_x1_bf, P.bf(nm.out_w), BF16)
    do = scatter_rows_into_zeros(do_t, idx, B * S)
    dqkv = torch.zeros_like(qkv)                       # dq outside the query window is exactly zero
    delta = torch.empty_like(lse)
    L.call("attn_window_bwd", qkv, o, do, lse, delta, dqkv, B, S, H, d, q0, qn, float(d ** -0.5))
    _linear_bwd(P, nm.qkv_w, nm.qkv_b, dqkv, a1)
    da1 = lin_dgrad(dqkv, P.bf(nm.qkv_w), BF16)
    d_res = scatter_rows_into_zeros(d_x1, idx, B * S)  # the residual path reaches the window rows only
    need1 = P.need(nm.ln1 + ".weight") or P.need(nm.ln1 + ".bias")
    d_z, d_z_bf = ln_bwd(da1, z, mu1, rs1, P[nm.ln1 + ".weight"], res1=d_res,
                         dw=P.gbuf(nm.ln1 + ".weight") if need1 else None, db=P.gbuf(nm.ln1 + ".bias") if need1 else None)
    return d_z, d_z_bf


# --------------------------------------------------------------------------------------------------
# video tower
# --------------------------------------------------------------------------------------------------
def video_param_names(cfg, prefix=""):
    v = prefix
    names = [v + "class_embedding", v + "positional_embedding", v + "temporal_embedding", v + "conv1.weight",
             v + "ln_pre.weight", v + "ln_pre.bias"]
    for i in range(cfg.layers):
        p = f"{v}transformer.resblocks.{i}."
        for a in ("timeattn", "attn"):
            names += [p + a + ".qkv.weight", p + a + ".qkv.bias", p + a + ".proj.weight", p + a + ".proj.bias"]
        for l in ("ln_3", "ln_1", "ln_2"):
            names += [p + l + ".weight", p + l + ".bias"]
        names += [p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", p + "mlp.c_proj.weight", p + "mlp.c_proj.bias"]
    names += [v + "ln_post.weight", v + "ln_post.bias", v + "proj"]
    return names


def video_forward(P, video, keep_ind, cfg):
    """VisionTransformer.forward -> vtok [B, N, E] fp32 (all tokens, ln_post + proj) and the saved activations."""
    if video.dim() == 4:
        video = video.unsqueeze(1)
    B, T = video.shape[0], video.shape[1]
    R, p, D, H = cfg.resolution, cfg.patch, cfg.width, cfg.heads
    n = keep_ind.shape[1]
    if n != cfg.kept_per_frame:
        raise ValueError(f"keep_ind keeps {n} patches/frame but the model expects {cfg.kept_per_frame}")
    if T > P["temporal_embedding"].shape[0]:
        raise ValueError(f"{T} frames exceed num_frames={P['temporal_embedding'].shape[0]}")
    N = 1 + T * n
    K = 3 * p * p
    as_u8 = video.dtype == torch.uint8      # uint8 crops: x/255 and (x-mean)/std of the CPU transform run inside the gather kernel
    video = video.contiguous() if as_u8 else video.contiguous().float()
    keep_ind = keep_ind.to(device=video.device, dtype=torch.int64).contiguous()
    tok = _empty((B * T * n, D), F32, video)
    if as_u8 and p % 4 != 0:
        raise NotImplementedError("uint8 clips are supported for patch sizes that are a multiple of 4 (B/16, B/32)")
    if p % 4 == 0:
        cols = _empty((B * T * n, K), BF16, video)
        if as_u8:
            import ctypes
            mean = (ctypes.c_float * 3)(*getattr(cfg, "input_mean", (0.485, 0.456, 0.406)))
            std = (ctypes.c_float * 3)(*getattr(cfg, "input_std", (0.229, 0.224, 0.225)))
            L.call("patch_gather_u8", video, keep_ind, cols, B, T, R, p, n, mean, std)
        else:
            L.call("patch_gather", video, keep_ind, cols, B, T, R, p, n)
        w_bf = P.bf("conv1.weight").view(D, K)
        L.gemm(cols, w_bf, tok, M=B * T * n, N=D, K=K, lda=K, ldb=K)
    else:
        # H/14: 3*14*14 = 588 bf16 per im2col row is not a whole number of 16-byte units -> both GEMM operands are laid out at a
        # row pitch of Kp (zero tail).  The padded weight copy is re-cast every step (1.5 MB) because the fused optimizer updates
        # conv1.weight in place without touching its version counter.
        Kp = (K + 7) // 8 * 8
        cols = _empty((B * T * n, Kp), BF16, video)
        L.call("patch_gather_ld", video, keep_ind, cols, B, T, R, p, n, Kp)
        w_bf = _empty((D, Kp), BF16, video)
        L.call("cast_bf16_pad", P["conv1.weight"].detach().contiguous(), w_bf, D, K, Kp)
        L.gemm(cols, w_bf, tok, M=B * T * n, N=D, K=Kp, lda=Kp, ldb=Kp)
    x0 = _empty((B * N, D), F32, video)
    L.call("video_assemble", tok, P["class_embedding"], P["positional_embedding"], P["temporal_embedding"], keep_ind, x0, B, T, n, D)
    x, mu0, rs0 = ln_fwd(x0, P["ln_pre.weight"], P["ln_pre.bias"], cfg.ln_eps, out_dtype=F32)
    blocks = []
    for i in range(cfg.layers):
        x, sv = st_block_fwd(P, f"transformer.resblocks.{i}.", x, B, N, T, n, H, cfg.act, cfg.ln_eps)
        blocks.append(sv)
    if cfg.post_mode == "h14":
        # video_encoder_ViT_H_14.py:472-484: pooled = ln_post(x[:, 0]) @ proj ; tokens = x[:, 1:] @ proj (no ln_post).  One GEMM over
        # all B*N rows whose operand holds the normalised CLS rows and the raw patch rows; row 0 of vtok is `pooled`.
        cls_idx = (torch.arange(B, device=x.device, dtype=torch.int64) * N).contiguous()
        xc = _empty((B, D), F32, x)
        L.call("gather_rows", x, cls_idx, xc, B, D)
        ac, mup, rsp = ln_fwd(xc, P["ln_post.weight"], P["ln_post.bias"], cfg.ln_eps)
        a = cast_bf16(x)
        L.call("scatter_rows", ac.view(F32), cls_idx, a.view(F32), B, D // 2, 0)
        x_last = xc
    else:
        cls_idx = None
        a, mup, rsp = ln_fwd(x, P["ln_post.weight"], P["ln_post.bias"], cfg.ln_eps)
        x_last = x
    vtok = mat_fwd(a, P.bf("proj"))
    saved = dict(B=B, T=T, n=n, N=N, keep=keep_ind, cols=cols, x0=x0, mu0=mu0, rs0=rs0, blocks=blocks, x_last=x_last, a=a, mup=mup,
                 rsp=rsp, cls_idx=cls_idx)
    return vtok.view(B, N, -1), saved


def video_backward(P, saved, d_vtok, cfg):
    B, T, n, N = saved["B"], saved["T"], saved["n"], saved["N"]
    D, H = cfg.width, cfg.heads
    E = d_vtok.shape[-1]
    d_vtok_bf = cast_bf16(d_vtok.reshape(B * N, E).contiguous())
    if P.need("proj"):
        mat_wgrad(saved["a"], d_vtok_bf, P.gbuf("proj"))
    da = mat_dgrad(d_vtok_bf, P.bf("proj"), F32)
    needp = P.need("ln_post.weight") or P.need("ln_post.bias")
    last_bias = _bias_buf(P, f"transformer.resblocks.{cfg.layers - 1}.mlp.c_proj.bias") if cfg.layers > 0 else None
    if cfg.post_mode == "h14":
        cls_idx = saved["cls_idx"]
        dac = _empty((B, D), F32, da)
        L.call("gather_rows", da, cls_idx, dac, B, D)
        d_xc, _ = ln_bwd(dac, saved["x_last"], saved["mup"], saved["rsp"], P["ln_post.weight"], want_bf16=False,
                         dw=P.gbuf("ln_post.weight") if needp else None, db=P.gbuf("ln_post.bias") if needp else None)
        L.call("scatter_rows", d_xc, cls_idx, da, B, D, 0)         # patch rows: d_x = da ; CLS rows: through ln_post
        d_x, d_x_bf = da, cast_bf16(da)
        if last_bias is not None:
            colsum(d_x_bf, last_bias)
    else:
        d_x, d_x_bf = ln_bwd(da, saved["x_last"], saved["mup"], saved["rsp"], P["ln_post.weight"],
                             dw=P.gbuf("ln_post.weight") if needp else None, db=P.gbuf("ln_post.bias") if needp else None,
                             dxsum=last_bias)
    for i in reversed(range(cfg.layers)):
        prev_bias = f"transformer.resblocks.{i - 1}.mlp.c_proj.bias" if i > 0 else None
        d_x, d_x_bf = st_block_bwd(P, f"transformer.resblocks.{i}.", saved["blocks"][i], d_x, d_x_bf, B, N, T, n, H, cfg.act, prev_bias)
        saved["blocks"][i] = None
        _grad_ready("video_block", i)      # (block i's own c_proj bias was filled by block i+1's LayerNorm backward, earlier)
    need0 = P.need("ln_pre.weight") or P.need("ln_pre.bias")
    d_x0, _ = ln_bwd(d_x, saved["x0"], saved["mu0"], saved["rs0"], P["ln_pre.weight"], want_bf16=False,
                     dw=P.gbuf("ln_pre.weight") if need0 else None, db=P.gbuf("ln_pre.bias") if need0 else None)
    dtok = _empty((B * T * n, D), BF16, d_x0)
    L.call("video_assemble_bwd", d_x0, saved["keep"], P.gbuf("class_embedding"), P.gbuf("positional_embedding"),
           P.gbuf("temporal_embedding"), dtok, B, T, n, D)
    if P.need("conv1.weight"):
        Kp = saved["cols"].shape[1]
        K = 3 * cfg.patch * cfg.patch
        if Kp == K:
            lin_wgrad(dtok, saved["cols"], P.gbuf("conv1.weight").view(D, K))
        else:                                                     # padded im2col rows (H/14): only the first K columns are real
            L.gemm(dtok, saved["cols"], P.gbuf("conv1.weight").view(D, K), M=D, N=K, K=dtok.shape[0], lda=D, ldb=Kp, a_mn=True,
                   b_mn=True, accumulate=True)


# set by trainer.TrainStep (gradient all-reduce overlapped with the backward): GRAD_READY(tag) is called, on the stream that produced
# them, whenever a set of parameter gradients has become final: ("video_block", i) after block i of the video tower's backward (blocks are
# walked from the last to the first), ("video_rest",) at the end of the video tower's backward, ("text",) / ("sort",) at the end of
# those towers' backward passes.
GRAD_READY = None


def _grad_ready(*tag):
    if GRAD_READY is not None:
        GRAD_READY(tag)


class _VideoTowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, names, video, keep_ind, *params):
        P = ParamView(names, params, [False] * len(names))
        vtok, saved = video_forward(P, video, keep_ind, cfg)
        ctx.cfg, ctx.names, ctx.saved, ctx.params = cfg, names, saved, params
        return vtok

    @staticmethod
    def backward(ctx, d_vtok):
        P = ParamView(ctx.names, ctx.params, ctx.needs_input_grad[4:])
        video_backward(P, ctx.saved, d_vtok, ctx.cfg)
        ctx.saved = None
        grads = (None, None, None, None) + P.grad_tuple()
        _grad_ready("video_rest")
        return grads


def video_tower(cfg, named_params, video, keep_ind):
    """named_params: ordered {local name (no 'video_model.' prefix): Parameter}"""
    names = list(named_params.keys())
    return _VideoTowerFn.apply(cfg, names, video, keep_ind, *named_params.values())


# --------------------------------------------------------------------------------------------------
# text tower
# --------------------------------------------------------------------------------------------------
def text_forward(P, tokens, cfg):
    """compute_text -> [n_txt, E] fp32 (ln_final + EOT pooling + text_projection)."""
    n_txt, Lc = tokens.shape
    W, H = cfg.text_width, cfg.text_heads
    is64 = int(tokens.dtype == torch.int64)
    tokens = tokens.contiguous()
    x = _empty((n_txt * Lc, W), F32, P["text_positional_embedding"])
    L.call("text_embed", tokens, is64, P["text_token_embedding.weight"], P["text_positional_embedding"], x, n_txt, Lc, W)
    eot = torch.empty(n_txt, dtype=torch.int64, device=x.device)
    L.call("argmax_rows", tokens, is64, eot, n_txt, Lc)
    blocks = []
    for i in range(cfg.text_layers):
        x, sv = block_fwd(P, clip_block_names(f"text_model.resblocks.{i}."), x, n_txt, Lc, H, cfg.text_act, cfg.ln_eps, True)
        blocks.append(sv)
    xe = _empty((n_txt, W), F32, x)
    L.call("gather_rows", x, eot, xe, n_txt, W)
    a, mu, rs = ln_fwd(xe, P["text_ln_final.weight"], P["text_ln_final.bias"], cfg.ln_eps)
    t = mat_fwd(a, P.bf("text_projection"))
    saved = dict(tokens=tokens, is64=is64, eot=eot, blocks=blocks, xe=xe, a=a, mu=mu, rs=rs, n_txt=n_txt, Lc=Lc)
    return t, saved


def text_backward(P, saved, d_t, cfg):
    n_txt, Lc = saved["n_txt"], saved["Lc"]
    W, H = cfg.text_width, cfg.text_heads
    d_t_bf = cast_bf16(d_t.contiguous())
    if P.need("text_projection"):
        mat_wgrad(saved["a"], d_t_bf, P.gbuf("text_projection"))
    da = mat_dgrad(d_t_bf, P.bf("text_projection"), F32)
    needf = P.need("text_ln_final.weight") or P.need("text_ln_final.bias")
    d_xe, _ = ln_bwd(da, saved["xe"], saved["mu"], saved["rs"], P["text_ln_final.weight"], want_bf16=False,
                     dw=P.gbuf("text_ln_final.weight") if needf else None, db=P.gbuf("text_ln_final.bias") if needf else None)
    d_x = _zeros((n_txt * Lc, W), F32, d_xe)
    L.call("scatter_rows", d_xe, saved["eot"], d_x, n_txt, W, 0)
    d_x_bf = cast_bf16(d_x)
    for i in reversed(range(cfg.text_layers)):
        d_x, d_x_bf = block_bwd(P, clip_block_names(f"text_model.resblocks.{i}."), saved["blocks"][i], d_x, d_x_bf, n_txt, Lc, H,
                                cfg.text_act, True)
        saved["blocks"][i] = None
    need_tab = P.need("text_token_embedding.weight")
    need_pos = P.need("text_positional_embedding")
    if need_tab or need_pos:
        L.call("text_embed_bwd", d_x, saved["tokens"], saved["is64"],
               P.gbuf("text_token_embedding.weight") if need_tab else None,
               P.gbuf("text_positional_embedding") if need_pos else None, n_txt, Lc, W)


class _TextTowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, names, tokens, *params):
        P = ParamView(names, params, [False] * len(names))
        t, saved = text_forward(P, tokens, cfg)
        ctx.cfg, ctx.names, ctx.saved, ctx.params = cfg, names, saved, params
        return t

    @staticmethod
    def backward(ctx, d_t):
        P = ParamView(ctx.names, ctx.params, ctx.needs_input_grad[3:])
        text_backward(P, ctx.saved, d_t, ctx.cfg)
        ctx.saved = None
        grads = (None, None, None) + P.grad_tuple()
        _grad_ready("text")
        return grads


def text_tower(cfg, named_params, tokens):
    names = list(named_params.keys())
    return _TextTowerFn.apply(cfg, names, tokens, *named_params.values())


# --------------------------------------------------------------------------------------------------
# sort head
# --------------------------------------------------------------------------------------------------
def sort_forward(P, text, vtok, cfg):
    """SortTransformer.forward.  text [n_trans*B, E] fp32 (clip-major rows tr*B+b, treated as constant), vtok [B, N, E] fp32
    -> logits [B, n_trans, n_trans] fp32."""
    B, N, E = vtok.shape
    nt = text.shape[0] // B
    S = N + nt
    H = cfg.sort_heads
    vtok = vtok.contiguous()
    text = text.contiguous()
    z = _empty((B * S, E), F32, vtok)
    L.call("sort_concat", vtok, text, P["type_embed"], z, B, N, nt, E)
    idx = (torch.arange(B, device=z.device, dtype=torch.int64)[:, None] * S + N +
           torch.arange(nt, device=z.device, dtype=torch.int64)[None, :]).reshape(-1).contiguous()
    blocks = []
    for i in range(cfg.sort_depth - 1):
        z, sv = block_fwd(P, sort_block_names(f"blocks.{i}."), z, B, S, H, "gelu", cfg.sort_ln_eps, False)
        blocks.append(sv)
    # the last block: transcript rows only
    zt, sv = window_block_fwd(P, sort_block_names(f"blocks.{cfg.sort_depth - 1}."), z, idx, B, S, N, nt, H, "gelu", cfg.sort_ln_eps)
    blocks.append(sv)
    y, mu, rs = ln_fwd(zt, P["norm.weight"], P["norm.bias"], cfg.sort_ln_eps, out_dtype=F32)
    C = P["head.weight"].shape[0]
    logits = _empty((B * nt, C), F32, z)
    L.call("small_linear_fwd", y, P["head.weight"], P["head.bias"], logits, B * nt, E, C)
    saved = dict(B=B, N=N, nt=nt, S=S, E=E, C=C, blocks=blocks, idx=idx, zt=zt, y=y, mu=mu, rs=rs)
    return logits.view(B, nt, C), saved


def sort_backward(P, saved, d_logits, cfg):
    B, N, nt, S, E, C = saved["B"], saved["N"], saved["nt"], saved["S"], saved["E"], saved["C"]
    H = cfg.sort_heads
    d_logits = d_logits.reshape(B * nt, C).contiguous()
    dy = _empty((B * nt, E), F32, d_logits)
    L.call("small_linear_bwd", d_logits, saved["y"], P["head.weight"], dy, P.gbuf("head.weight"), P.gbuf("head.bias"), B * nt, E, C)
    d_zt, _ = ln_bwd(dy, saved["zt"], saved["mu"], saved["rs"], P["norm.weight"], want_bf16=False,
                     dw=P.gbuf("norm.weight"), db=P.gbuf("norm.bias"))
    last = cfg.sort_depth - 1
    d_z, d_z_bf = window_block_bwd(P, sort_block_names(f"blocks.{last}."), saved["blocks"][last], d_zt, saved["idx"], B, S, N, nt, H, "gelu")
    saved["blocks"][last] = None
    for i in reversed(range(last)):
        d_z, d_z_bf = block_bwd(P, sort_block_names(f"blocks.{i}."), saved["blocks"][i], d_z, d_z_bf, B, S, H, "gelu", False)
        saved["blocks"][i] = None
    d_vtok = _empty((B, N, E), F32, d_z)
    L.call("sort_concat_bwd", d_z, d_vtok, P.gbuf("type_embed"), B, N, nt, E)
    return d_vtok


class _SortHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, names, text, vtok, *params):
        P = ParamView(names, params, [False] * len(names))
        logits, saved = sort_forward(P, text, vtok, cfg)
        ctx.cfg, ctx.names, ctx.saved, ctx.params = cfg, names, saved, params
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        P = ParamView(ctx.names, ctx.params, ctx.needs_input_grad[4:])
        d_vtok = sort_backward(P, ctx.saved, d_logits, ctx.cfg)
        ctx.saved = None
        grads = (None, None, None, d_vtok if ctx.needs_input_grad[3] else None) + P.grad_tuple()
        _grad_ready("sort")
        return grads


def sort_head(cfg, named_params, text_detached, vtok):
    names = list(named_params.keys())
    return _SortHeadFn.apply(cfg, names, text_detached, vtok, *named_params.values())


# --------------------------------------------------------------------------------------------------
# transcript mean, similarity matrix, losses
# --------------------------------------------------------------------------------------------------
class _GroupMeanFn(torch.autograd.Function):
    """t [n_trans*B, E] (clip-major) -> mean over transcripts [B, E]   (model_dist_TVTSv2_ViT_B_16.py:74-76)"""

    @staticmethod
    def forward(ctx, t, nt):
        t = t.contiguous()
        B = t.shape[0] // nt
        E = t.shape[1]
        out = _empty((B, E), F32, t)
        L.call("group_mean", t, out, nt, B, E)
        ctx.nt, ctx.B, ctx.E = nt, B, E
        return out

    @staticmethod
    def backward(ctx, dout):
        dt = _empty((ctx.nt * ctx.B, ctx.E), F32, dout)
        L.call("group_mean_bwd", dout.contiguous(), dt, None, ctx.nt, ctx.B, ctx.E)
        return dt, None


def group_mean(t, nt):
    return _GroupMeanFn.apply(t, nt)


class _SimMatrixFn(torch.autograd.Function):
    """sim_matrix(a, b, eps): rows L2-normalised with a clamp, a_n @ b_n^T   (model_dist_TVTSv2_ViT_B_16.py:119-127)"""

    @staticmethod
    def forward(ctx, a, b, eps):
        a, b = a.contiguous().float(), b.contiguous().float()
        Ra, E = a.shape
        Rb = b.shape[0]
        an, bn = torch.empty_like(a), torch.empty_like(b)
        na, nb = _empty((Ra,), F32, a), _empty((Rb,), F32, a)
        L.call("normalize_rows", a, an, na, Ra, E, float(eps))
        L.call("normalize_rows", b, bn, nb, Rb, E, float(eps))
        S = _empty((Ra, Rb), F32, a)
        L.call("sim_matrix", an, bn, S, Ra, Rb, E, 1.0)
        ctx.saved = (an, bn, na, nb)
        ctx.eps = float(eps)
        return S

    @staticmethod
    def backward(ctx, G):
        an, bn, na, nb = ctx.saved
        Ra, E = an.shape
        Rb = bn.shape[0]
        G = G.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(an)
            L.call("sim_matrix_bwd", G, an, bn, na, da, Ra, Rb, E, 0, Ra, 0, 1.0, ctx.eps)
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(bn)
            L.call("sim_matrix_bwd", G, bn, an, nb, db, Rb, Ra, E, 0, Rb, 1, 1.0, ctx.eps)
        return da, db, None


def sim_matrix(a, b, eps=1e-8):
    return _SimMatrixFn.apply(a, b, eps)


class _NormSoftmaxLossFn(torch.autograd.Function):
    """NormSoftmaxLoss.forward  (v2/model/loss.py:13-25)"""

    @staticmethod
    def forward(ctx, S, temperature):
        S = S.contiguous().float()
        Bg = S.shape[0]
        if S.shape[1] != Bg:
            raise ValueError("NormSoftmaxLoss expects a square similarity matrix")
        lse_r, lse_c = _empty((Bg,), F32, S), _empty((Bg,), F32, S)
        loss = _empty((), F32, S)
        L.call("nsl_fwd", S, lse_r, lse_c, loss, Bg, float(temperature))
        ctx.saved = (S, lse_r, lse_c)
        ctx.temperature = float(temperature)
        return loss

    @staticmethod
    def backward(ctx, gout):
        S, lse_r, lse_c = ctx.saved
        G = torch.empty_like(S)
        L.call("nsl_bwd", S, lse_r, lse_c, gout.contiguous().float(), G, S.shape[0], ctx.temperature)
        return G, None


def norm_softmax_loss(S, temperature=0.05):
    return _NormSoftmaxLossFn.apply(S, temperature)


class _SortCEFn(torch.autograd.Function):
    """2 * CrossEntropyLoss()(pred.reshape(-1, C), labels.reshape(-1))   (v2/trainer/trainer.py:487-492)"""

    @staticmethod
    def forward(ctx, pred, labels, weight):
        C = pred.shape[-1]
        x = pred.reshape(-1, C).contiguous().float()
        y = labels.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
        loss = _empty((), F32, x)
        L.call("sort_ce", x, y, None, loss, None, x.shape[0], C, float(weight))
        ctx.saved = (x, y)
        ctx.weight = float(weight)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, gout):
        x, y = ctx.saved
        dx = torch.empty_like(x)
        L.call("sort_ce", x, y, gout.contiguous().float(), None, dx, x.shape[0], x.shape[1], ctx.weight)
        return dx.view(ctx.shape), None, None


def sort_ce(pred, labels, weight=2.0):
    return _SortCEFn.apply(pred, labels, weight)


class _FusedLossesFn(torch.autograd.Function):
    """Everything between the towers and their backward passes as ONE kernel launch (csrc/loss_fused.cu): the embedding all-gather
    (one NCCL collective, AllGather_multi semantics: v2/trainer/trainer.py:41-57), sim_matrix + NormSoftmaxLoss on the gathered
    embeddings with the gradient of the LOCAL rows, and 2 * CE of the sort logits.  Returns (loss1, loss2)."""

    @staticmethod
    def forward(ctx, video, text, pred, labels, temperature, eps, weight, gather):
        video, text = video.contiguous().float(), text.contiguous().float()
        B, E_ = video.shape
        video_all, text_all, row0 = gather(video, text)
        Bg = video_all.shape[0]
        loss1 = _empty((), F32, video)
        loss2 = torch.zeros((), dtype=F32, device=video.device)
        dv, dt = torch.empty_like(video), torch.empty_like(text)
        if pred is not None:
            C = pred.shape[-1]
            x = pred.reshape(-1, C).contiguous().float()
            y = labels.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
            dx = torch.empty_like(x)
            L.call("contrastive_sortce_fused", video_all, text_all, Bg, E_, row0, B, float(temperature), float(eps), x, y, x.shape[0], C,
                   float(weight), loss1, loss2, dv, dt, dx)
            ctx.pred_shape = pred.shape
        else:
            dx = None
            L.call("contrastive_sortce_fused", video_all, text_all, Bg, E_, row0, B, float(temperature), float(eps), None, None, 0, 0,
                   float(weight), loss1, None, dv, dt, None)
            ctx.pred_shape = None
        ctx.saved = (dv, dt, dx)
        ctx.mark_non_differentiable(loss2) if pred is None else None
        return loss1, loss2

    @staticmethod
    def backward(ctx, g1, g2):
        dv, dt, dx = ctx.saved
        ctx.saved = None
        dpred = (dx * g2).view(ctx.pred_shape) if dx is not None else None
        return dv * g1, dt * g1, dpred, None, None, None, None, None


def fused_losses_supported(Bg, E_):
    return bool(L.lib().tvts_contrastive_sortce_fused_supported(int(Bg), int(E_)))


def fused_losses(video, text, pred, labels, temperature, gather, eps=1e-8, weight=2.0):
    return _FusedLossesFn.apply(video, text, pred, labels, temperature, eps, weight, gather)
