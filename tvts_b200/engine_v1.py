"""Host-side orchestration of the TVTS v1 hot path (BASELINE.json configs[4]) on the same C-ABI kernels as engine.py.

Reference semantics restated here (paths relative to /root/reference):
  video tower   v1/model/video_encoder.py:78-99 (Conv3d tubelet embed), :59-75 (pre-LN block, joint attention :30-56),
                :178-217 (forward_features: embeddings added before the per-tube mask gather, final norm)
  text tower    DistilBERT (`AutoModel.from_pretrained('distilbert-base-uncased')`, v1/model/model_dist_TVTS.py:33,124-126) -- an
                un-vendored `transformers` dependency, restated from its published algorithm (oracle/tvts_oracle.py:distilbert_forward
                is the pinned restatement): POST-LN blocks, separate q/k/v Linears, key-padding mask, erf GELU, LayerNorm eps 1e-12
  heads         txt_proj = Sequential(ReLU, Linear), vid_proj = Sequential(Linear)   v1/model/model_dist_TVTS.py:64-74
Like engine.py: one autograd node per tower with a hand-written backward; fp32 residual stream / LayerNorm / softmax statistics, bf16
GEMM operands.  The v1 video block (timm-style pre-LN block, qkv bias, erf GELU, eps 1e-6) is exactly the sort head's block, so it reuses
engine.block_fwd / block_bwd.
"""
import torch

from . import _lib as L
from . import engine as E
from .engine import BF16, F32, ParamView, _empty, _zeros

LN_EPS_V1 = 1e-6
LN_EPS_BERT = 1e-12


# --------------------------------------------------------------------------------------------------
# video tower
# --------------------------------------------------------------------------------------------------
def video_param_names(depth):
    names = ["cls_token", "pos_embed", "temporal_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    for i in range(depth):
        p = f"blocks.{i}."
        names += [p + "norm1.weight", p + "norm1.bias", p + "attn.qkv.weight", p + "attn.qkv.bias", p + "attn.proj.weight",
                  p + "attn.proj.bias", p + "norm2.weight", p + "norm2.bias", p + "mlp.fc1.weight", p + "mlp.fc1.bias",
                  p + "mlp.fc2.weight", p + "mlp.fc2.bias"]
    return names + ["norm.weight", "norm.bias"]


def video_forward(P, video, keep_ind, cfg):
    """VisionTransformer.forward_features -> [B, N, D] fp32 (all tokens after the final norm), N = 1 + (T/2) * n."""
    B, T = video.shape[0], video.shape[1]
    R, p, D, H = video.shape[-1], cfg.patch, cfg.width, cfg.heads
    if T % 2:
        raise ValueError(f"v1 tubelets span 2 frames: T={T} must be even")
    nt = T // 2
    if keep_ind.dim() != 3 or keep_ind.shape[1] < nt:
        raise ValueError(f"keep_ind must be [B, >= T/2, n] (one subset per tube), got {tuple(keep_ind.shape)}")
    if nt > P["temporal_embed"].shape[1]:
        raise ValueError(f"{nt} tubes exceed the temporal table ({P['temporal_embed'].shape[1]})")
    n = keep_ind.shape[2]
    N = 1 + nt * n
    K = 3 * 2 * p * p
    video = video.contiguous().float()
    keep = keep_ind[:, :nt].to(device=video.device, dtype=torch.int64).contiguous()
    cols = _empty((B * nt * n, K), BF16, video)
    L.call("tubelet_gather", video, keep, cols, B, T, R, p, n)
    tok = _empty((B * nt * n, D), F32, video)
    L.gemm(cols, P.bf("patch_embed.proj.weight").view(D, K), tok, M=B * nt * n, N=D, K=K, lda=K, ldb=K, bias=P["patch_embed.proj.bias"])
    x = _empty((B * N, D), F32, video)
    L.call("video_assemble_tube", tok, P["cls_token"], P["pos_embed"], P["temporal_embed"], keep, x, B, nt, n, D)
    blocks = []
    for i in range(cfg.layers):
        x, sv = E.block_fwd(P, E.sort_block_names(f"blocks.{i}."), x, B, N, H, "gelu", LN_EPS_V1, False)
        blocks.append(sv)
    y, mu, rs = E.ln_fwd(x, P["norm.weight"], P["norm.bias"], LN_EPS_V1, out_dtype=F32)
    saved = dict(B=B, nt=nt, n=n, N=N, keep=keep, cols=cols, blocks=blocks, x_last=x, mu=mu, rs=rs)
    return y.view(B, N, D), saved


def video_backward(P, saved, d_y, cfg):
    B, nt, n, N = saved["B"], saved["nt"], saved["n"], saved["N"]
    D, H = cfg.width, cfg.heads
    d_y = d_y.reshape(B * N, D).contiguous()
    need = P.need("norm.weight") or P.need("norm.bias")
    d_x, d_x_bf = E.ln_bwd(d_y, saved["x_last"], saved["mu"], saved["rs"], P["norm.weight"],
                           dw=P.gbuf("norm.weight") if need else None, db=P.gbuf("norm.bias") if need else None)
    for i in reversed(range(cfg.layers)):
        d_x, d_x_bf = E.block_bwd(P, E.sort_block_names(f"blocks.{i}."), saved["blocks"][i], d_x, d_x_bf, B, N, H, "gelu", False)
        saved["blocks"][i] = None
    dtok = _empty((B * nt * n, D), BF16, d_x)
    L.call("video_assemble_tube_bwd", d_x, saved["keep"], P.gbuf("cls_token"), P.gbuf("pos_embed"), P.gbuf("temporal_embed"), dtok,
           B, nt, n, D)
    if P.need("patch_embed.proj.weight"):
        K = saved["cols"].shape[1]
        E.lin_wgrad(dtok, saved["cols"], P.gbuf("patch_embed.proj.weight").view(D, K))
    if P.need("patch_embed.proj.bias"):
        E.colsum(dtok, P.gbuf("patch_embed.proj.bias"))


class _VideoTowerV1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, names, video, keep_ind, *params):
        P = ParamView(names, params, [False] * len(names))
        y, saved = video_forward(P, video, keep_ind, cfg)
        ctx.cfg, ctx.names, ctx.saved, ctx.params = cfg, names, saved, params
        return y

    @staticmethod
    def backward(ctx, d_y):
        P = ParamView(ctx.names, ctx.params, ctx.needs_input_grad[4:])
        video_backward(P, ctx.saved, d_y, ctx.cfg)
        ctx.saved = None
        return (None, None, None, None) + P.grad_tuple()


def video_tower(cfg, named_params, video, keep_ind):
    return _VideoTowerV1Fn.apply(cfg, list(named_params.keys()), video, keep_ind, *named_params.values())


# --------------------------------------------------------------------------------------------------
# DistilBERT text tower -> [CLS] hidden state
# --------------------------------------------------------------------------------------------------
def distil_param_names(layers):
    names = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight", "embeddings.LayerNorm.weight",
             "embeddings.LayerNorm.bias"]
    for i in range(layers):
        p = f"transformer.layer.{i}."
        for lin in ("attention.q_lin", "attention.k_lin", "attention.v_lin", "attention.out_lin"):
            names += [p + lin + ".weight", p + lin + ".bias"]
        names += [p + "sa_layer_norm.weight", p + "sa_layer_norm.bias", p + "ffn.lin1.weight", p + "ffn.lin1.bias", p + "ffn.lin2.weight",
                  p + "ffn.lin2.bias", p + "output_layer_norm.weight", p + "output_layer_norm.bias"]
    return names


def _ln_post(P, name, s):
    """x = LayerNorm(s) of a post-LN block: the fp32 residual stream and its bf16 copy (the next GEMM operand)."""
    x, mu, rs = E.ln_fwd(s, P[name + ".weight"], P[name + ".bias"], LN_EPS_BERT, out_dtype=F32)
    return x, E.cast_bf16(x), mu, rs


def _ln_grads(P, name):
    need = P.need(name + ".weight") or P.need(name + ".bias")
    return (P.gbuf(name + ".weight"), P.gbuf(name + ".bias")) if need else (None, None)


def distil_forward(P, input_ids, attention_mask, cfg):
    """DistilBertModel(input_ids, attention_mask).last_hidden_state[:, 0] -> [n_txt, W] fp32."""
    n_txt, Lc = input_ids.shape
    W, H = cfg.text_width, cfg.text_heads
    d = W // H
    M = n_txt * Lc
    if Lc > P["embeddings.position_embeddings.weight"].shape[0]:
        raise ValueError(f"sequence length {Lc} exceeds max_position_embeddings")
    ids = input_ids.contiguous()
    is64 = int(ids.dtype == torch.int64)
    # the tokenizer pads on the right: the mask is a run of ones, so a per-sequence key count describes it
    klen = attention_mask.to(torch.int32).sum(1).to(torch.int32).contiguous()
    xe = _empty((M, W), F32, P["embeddings.LayerNorm.weight"])
    L.call("text_embed", ids, is64, P["embeddings.word_embeddings.weight"], P["embeddings.position_embeddings.weight"], xe, n_txt, Lc, W)
    x, x_bf, mu_e, rs_e = _ln_post(P, "embeddings.LayerNorm", xe)
    scale = float(d ** -0.5)
    layers = []
    for i in range(cfg.text_layers):
        p = f"transformer.layer.{i}."
        qkv = _empty((M, 3 * W), BF16, x)
        for j, lin in enumerate(("q_lin", "k_lin", "v_lin")):       # three Linears write the packed [M, 3, H, d] buffer the attention reads
            L.gemm(x_bf, P.bf(p + f"attention.{lin}.weight"), qkv[:, j * W:(j + 1) * W], M=M, N=W, K=W, lda=W, ldb=W, ldo=3 * W,
                   bias=P[p + f"attention.{lin}.bias"])
        o = _empty((M, W), BF16, x)
        lse = _empty((n_txt, H, Lc), F32, x)
        L.call("attn_padded_fwd", qkv, o, lse, klen, n_txt, Lc, H, d, scale)
        s1 = E.lin_fwd(o, P.bf(p + "attention.out_lin.weight"), P[p + "attention.out_lin.bias"], F32, residual=x)
        x1, x1_bf, mu1, rs1 = _ln_post(P, p + "sa_layer_norm", s1)
        g, h = E.lin_fwd(x1_bf, P.bf(p + "ffn.lin1.weight"), P[p + "ffn.lin1.bias"], BF16, act="gelu", want_pre=True)
        s2 = E.lin_fwd(g, P.bf(p + "ffn.lin2.weight"), P[p + "ffn.lin2.bias"], F32, residual=x1)
        layers.append((x_bf, qkv, o, lse, s1, mu1, rs1, x1_bf, h, g, s2))
        x, x_bf, mu2, rs2 = _ln_post(P, p + "output_layer_norm", s2)
        layers[-1] = layers[-1] + (mu2, rs2)
    cls_idx = (torch.arange(n_txt, device=x.device, dtype=torch.int64) * Lc).contiguous()
    cls = _empty((n_txt, W), F32, x)
    L.call("gather_rows", x, cls_idx, cls, n_txt, W)
    saved = dict(ids=ids, is64=is64, klen=klen, xe=xe, mu_e=mu_e, rs_e=rs_e, layers=layers, cls_idx=cls_idx, n_txt=n_txt, Lc=Lc, scale=scale)
    return cls, saved


def distil_backward(P, saved, d_cls, cfg):
    n_txt, Lc, klen, scale = saved["n_txt"], saved["Lc"], saved["klen"], saved["scale"]
    W, H = cfg.text_width, cfg.text_heads
    d = W // H
    M = n_txt * Lc
    d_x = _zeros((M, W), F32, d_cls)
    L.call("scatter_rows", d_cls.contiguous().float(), saved["cls_idx"], d_x, n_txt, W, 0)
    for i in reversed(range(cfg.text_layers)):
        p = f"transformer.layer.{i}."
        (x_bf, qkv, o, lse, s1, mu1, rs1, x1_bf, h, g, s2, mu2, rs2) = saved["layers"][i]
        saved["layers"][i] = None
        # x_out = LN(s2), s2 = lin2(gelu(lin1(x1))) + x1
        dw, db = _ln_grads(P, p + "output_layer_norm")
        d_s2, d_s2_bf = E.ln_bwd(d_x, s2, mu2, rs2, P[p + "output_layer_norm.weight"], dw=dw, db=db)
        E._linear_bwd(P, p + "ffn.lin2.weight", p + "ffn.lin2.bias", d_s2_bf, g)
        dh = E.lin_dgrad(d_s2_bf, P.bf(p + "ffn.lin2.weight"), BF16, dact="gelu", aux=h)
        E._linear_bwd(P, p + "ffn.lin1.weight", p + "ffn.lin1.bias", dh, x1_bf)
        w1 = P.bf(p + "ffn.lin1.weight")                                   # [4W, W]
        d_x1 = _empty((M, W), F32, d_x)
        L.gemm(dh, w1, d_x1, M=M, N=W, K=w1.shape[0], lda=w1.shape[0], ldb=W, b_mn=True, residual=d_s2)     # dgrad + the residual path
        # x1 = LN(s1), s1 = out_lin(attn) + x
        dw, db = _ln_grads(P, p + "sa_layer_norm")
        d_s1, d_s1_bf = E.ln_bwd(d_x1, s1, mu1, rs1, P[p + "sa_layer_norm.weight"], dw=dw, db=db)
        E._linear_bwd(P, p + "attention.out_lin.weight", p + "attention.out_lin.bias", d_s1_bf, o)
        do = E.lin_dgrad(d_s1_bf, P.bf(p + "attention.out_lin.weight"), BF16)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        L.call("attn_padded_bwd", qkv, o, do, lse, delta, dqkv, klen, n_txt, Lc, H, d, scale)
        d_x = _empty((M, W), F32, d_s1)
        for j, lin in enumerate(("q_lin", "k_lin", "v_lin")):
            dy = dqkv[:, j * W:(j + 1) * W]                               # [M, W] column block, row pitch 3W
            wn, bn = p + f"attention.{lin}.weight", p + f"attention.{lin}.bias"
            if P.need(wn):
                L.gemm(dy, x_bf, P.gbuf(wn), M=W, N=W, K=M, lda=3 * W, ldb=W, a_mn=True, b_mn=True, accumulate=True)
            if P.need(bn):
                L.call("colsum_bf16", dy, P.gbuf(bn), M, W, 3 * W)
            if j == 0:                                                    # d_x = d_s1 + sum_j dy_j W_j
                L.gemm(dy, P.bf(wn), d_x, M=M, N=W, K=W, lda=3 * W, ldb=W, b_mn=True, residual=d_s1)
            else:
                L.gemm(dy, P.bf(wn), d_x, M=M, N=W, K=W, lda=3 * W, ldb=W, b_mn=True, accumulate=True)
    dw, db = _ln_grads(P, "embeddings.LayerNorm")
    d_xe, _ = E.ln_bwd(d_x, saved["xe"], saved["mu_e"], saved["rs_e"], P["embeddings.LayerNorm.weight"], want_bf16=False, dw=dw, db=db)
    need_tab = P.need("embeddings.word_embeddings.weight")
    need_pos = P.need("embeddings.position_embeddings.weight")
    if need_tab or need_pos:
        L.call("text_embed_bwd", d_xe, saved["ids"], saved["is64"], P.gbuf("embeddings.word_embeddings.weight") if need_tab else None,
               P.gbuf("embeddings.position_embeddings.weight") if need_pos else None, n_txt, Lc, W)


class _DistilFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, names, input_ids, attention_mask, *params):
        P = ParamView(names, params, [False] * len(names))
        cls, saved = distil_forward(P, input_ids, attention_mask, cfg)
        ctx.cfg, ctx.names, ctx.saved, ctx.params = cfg, names, saved, params
        return cls

    @staticmethod
    def backward(ctx, d_cls):
        P = ParamView(ctx.names, ctx.params, ctx.needs_input_grad[4:])
        distil_backward(P, ctx.saved, d_cls, ctx.cfg)
        ctx.saved = None
        return (None, None, None, None) + P.grad_tuple()


def distil_cls(cfg, named_params, input_ids, attention_mask):
    return _DistilFn.apply(cfg, list(named_params.keys()), input_ids, attention_mask, *named_params.values())


# --------------------------------------------------------------------------------------------------
# projection heads:  y = [relu](x) W^T + b
# --------------------------------------------------------------------------------------------------
class _ProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, relu, x, w, b):
        x = x.contiguous().float()
        R, K = x.shape
        if relu:
            a = _empty((R, K), BF16, x)
            L.call("relu_bf16", x, a, R * K)
        else:
            a = E.cast_bf16(x)
        y = E.lin_fwd(a, E.WEIGHTS.get(w), b, F32)
        ctx.relu, ctx.saved = relu, (x, a, w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a, w, b = ctx.saved
        P = ParamView(["w", "b"], [w, b], ctx.needs_input_grad[2:4])
        dy_bf = E.cast_bf16(dy.contiguous().float())
        E._linear_bwd(P, "w", "b", dy_bf, a)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = E.lin_dgrad(dy_bf, E.WEIGHTS.get(w), F32)
            if ctx.relu:
                out = torch.empty_like(dx)
                L.call("relu_bwd", x, dx, out, dx.numel())
                dx = out
        ctx.saved = None
        return (None, dx) + P.grad_tuple()


def projection(x, weight, bias, relu=False):
    return _ProjFn.apply(relu, x, weight, bias)
