"""CPU oracle for the TVTSv2 pre-training hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A from-scratch functional restatement (torch CPU, fp32 or fp64, autograd for gradients) of the
reference's algorithm, operating directly on a reference-named state_dict.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Pinning: the reference ships no tests or golden vectors for this path, so this
restatement is pinned against outputs of the UNMODIFIED reference classes executed in the build
container (oracle/make_golden.py -> tests/golden/*.npz; checked by tests/test_oracle_golden.py).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    """v2/model/video_encoder_ViT_B_16.py:79-85 (fp32 LayerNorm); sort head uses eps=1e-6 (sort_transformer.py:100)."""
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * w + b


def quick_gelu(x):
    """v2/model/video_encoder_ViT_B_16.py:88-90"""
    return x * torch.sigmoid(1.702 * x)


def gelu_erf(x):
    """nn.GELU() default (erf) used by sort_transformer.py:17 and the H/14 towers."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _act(name):
    return quick_gelu if name == "quick_gelu" else gelu_erf


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def _softmax_av(q, k, v):
    """softmax(q k^T) v  over the last two dims (scale already folded into q)."""
    s = q @ k.transpose(-1, -2)
    p = torch.softmax(s, dim=-1)
    return p @ v


# --------------------------------------------------------------------------------------------------
# divided space-time attention (VarAttention)
# --------------------------------------------------------------------------------------------------
def var_attention(x, qkv_w, qkv_b, proj_w, proj_b, heads, T, n, mode):
    """v2/model/video_encoder_ViT_B_16.py:38-76.

    x: [B, 1+T*n, D].  mode 'time': patch (f, j) attends to [CLS ; (f', j) for all f'];
    mode 'space': patch (f, j) attends to [CLS ; (f, j') for all j'].  CLS attends to everything.
    """
    B, N, D = x.shape
    d = D // heads
    qkv = linear(x, qkv_w, qkv_b).reshape(B, N, 3, heads, d)
    q = qkv[:, :, 0].permute(0, 2, 1, 3) * (d ** -0.5)       # [B,h,N,d]  (:45 scale on q)
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)

    cls_out = _softmax_av(q[:, :, :1], k, v)                  # [B,h,1,d]   (:51)

    def grid(t):                                              # patch tokens as [B,h,T,n,d]
        return t[:, :, 1:].reshape(B, heads, T, n, d)

    qg, kg, vg = grid(q), grid(k), grid(v)
    if mode == "time":                                        # '(b n) f d': groups over the frame axis
        qg, kg, vg = (t.transpose(2, 3) for t in (qg, kg, vg))  # [B,h,n,T,d]
    G, L = qg.shape[2], qg.shape[3]
    cls_k = k[:, :, :1].unsqueeze(2).expand(B, heads, G, 1, d)  # CLS key/value prepended (:56-60)
    cls_v = v[:, :, :1].unsqueeze(2).expand(B, heads, G, 1, d)
    out = _softmax_av(qg, torch.cat([cls_k, kg], 3), torch.cat([cls_v, vg], 3))  # [B,h,G,L,d]
    if mode == "time":
        out = out.transpose(2, 3)
    out = out.reshape(B, heads, T * n, d)
    out = torch.cat([cls_out, out], 2)                        # [B,h,N,d]  (:69)
    out = out.permute(0, 2, 1, 3).reshape(B, N, D)            # merge heads (:72)
    return linear(out, proj_w, proj_b)


def st_block(x, sd, p, heads, T, n, act, eps):
    """ResidualSpaceTimeAttentionBlock.forward, v2/model/video_encoder_ViT_B_16.py:113-124."""
    g = lambda k: sd[p + k]
    t = var_attention(layer_norm(x, g("ln_3.weight"), g("ln_3.bias"), eps),
                      g("timeattn.qkv.weight"), g("timeattn.qkv.bias"),
                      g("timeattn.proj.weight"), g("timeattn.proj.bias"), heads, T, n, "time")
    time_res = x + t
    s = var_attention(layer_norm(time_res, g("ln_1.weight"), g("ln_1.bias"), eps),
                      g("attn.qkv.weight"), g("attn.qkv.bias"),
                      g("attn.proj.weight"), g("attn.proj.bias"), heads, T, n, "space")
    space_res = x + s                                          # residual is x, NOT time_res (:121)
    h = layer_norm(space_res, g("ln_2.weight"), g("ln_2.bias"), eps)
    h = linear(_act(act)(linear(h, g("mlp.c_fc.weight"), g("mlp.c_fc.bias"))),
               g("mlp.c_proj.weight"), g("mlp.c_proj.bias"))
    return space_res + h


# --------------------------------------------------------------------------------------------------
# video tower
# --------------------------------------------------------------------------------------------------
def patchify(video, p):
    """Conv2d(k=s=p, no bias) as im2col: [B,T,3,H,W] -> [B,T,P,3*p*p] with (c,u,v) column order
    (matches conv1.weight.reshape(D,-1)); patches row-major.  video_encoder_ViT_B_16.py:180-184."""
    B, T, C, H, W = video.shape
    g = H // p
    x = video.reshape(B, T, C, g, p, g, p).permute(0, 1, 3, 5, 2, 4, 6)
    return x.reshape(B, T, g * g, C * p * p)


def video_embed(video, keep_ind, sd, cfg, prefix="video_model."):
    """Patch-embed + CLS + positional/temporal embedding + tube-mask gather (before ln_pre).
    video_encoder_ViT_B_16.py:176-216.  keep_ind [B,n] int64 is shared by all frames of a sample."""
    if video.dim() == 4:
        video = video.unsqueeze(1)
    B, T = video.shape[:2]
    D = cfg.width
    w = sd[prefix + "conv1.weight"].reshape(D, -1)
    pos = sd[prefix + "positional_embedding"]
    tem = sd[prefix + "temporal_embedding"]
    cols = patchify(video, cfg.patch)                                   # [B,T,P,K]
    idx = keep_ind.to(torch.long)[:, None, :, None].expand(B, T, keep_ind.shape[1], cols.shape[-1])
    kept = torch.gather(cols, 2, idx)                                   # gather BEFORE the GEMM (per-patch independent)
    tok = kept @ w.t()                                                  # [B,T,n,D]
    tok = tok + pos[1:][keep_ind.to(torch.long)][:, None] + tem[:T][None, :, None, :]
    cls = (sd[prefix + "class_embedding"] + pos[0]).expand(B, 1, D)
    return torch.cat([cls, tok.reshape(B, -1, D)], 1)                   # [B, 1+T*n, D]


def video_tower(video, keep_ind, sd, cfg, prefix="video_model."):
    """VisionTransformer.forward, v2/model/video_encoder_ViT_B_16.py:176-235 -> [B,N,E].
    cfg.post_mode == 'h14' (v2/model/video_encoder_ViT_H_14.py:419-484, exact-GELU blocks via cfg.act): returns
    (pooled [B,E] = ln_post(x[:,0]) @ proj, tokens [B,N-1,E] = x[:,1:] @ proj -- no ln_post, no CLS; :472-482)."""
    T = 1 if video.dim() == 4 else video.shape[1]
    n = keep_ind.shape[1]
    assert n == cfg.kept_per_frame, (n, cfg.kept_per_frame)            # :220 time_n must equal kept count
    x = video_embed(video, keep_ind, sd, cfg, prefix)
    x = layer_norm(x, sd[prefix + "ln_pre.weight"], sd[prefix + "ln_pre.bias"], cfg.ln_eps)
    for i in range(cfg.layers):
        x = st_block(x, sd, f"{prefix}transformer.resblocks.{i}.", cfg.heads, T, n, cfg.act, cfg.ln_eps)
    if getattr(cfg, "post_mode", "all") == "h14":
        pooled = layer_norm(x[:, 0], sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"], cfg.ln_eps) @ sd[prefix + "proj"]
        return pooled, x[:, 1:] @ sd[prefix + "proj"]
    x = layer_norm(x, sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"], cfg.ln_eps)
    return x @ sd[prefix + "proj"]


# --------------------------------------------------------------------------------------------------
# text tower (CLIP)
# --------------------------------------------------------------------------------------------------
def text_tower(tokens, sd, cfg):
    """TVTSv2_B_16.compute_text, model_dist_TVTSv2_ViT_B_16.py:97-111 with the CLIP transformer
    (v2/CLIP/clip/model.py:171-203, causal mask :330-336).  tokens [n_txt, ctx] int -> [n_txt, E]."""
    W, h = cfg.text_width, cfg.text_heads
    d = W // h
    n_txt, L = tokens.shape
    x = sd["text_token_embedding.weight"][tokens.to(torch.long)] + sd["text_positional_embedding"]
    causal = torch.full((L, L), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
    act = _act(cfg.text_act)
    for i in range(cfg.text_layers):
        p = f"text_model.resblocks.{i}."
        y = layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], cfg.ln_eps)
        qkv = linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]).reshape(n_txt, L, 3, h, d)
        q = qkv[:, :, 0].transpose(1, 2) * (d ** -0.5)
        k = qkv[:, :, 1].transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        s = q @ k.transpose(-1, -2) + causal
        o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_txt, L, W)
        x = x + linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], cfg.ln_eps)
        x = x + linear(act(linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])),
                       sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    eot = tokens.to(torch.long).argmax(-1)                               # EOT has the highest id (:107-108)
    x = x[torch.arange(n_txt, device=x.device), eot]                                      # LN is per-row: gather first
    x = layer_norm(x, sd["text_ln_final.weight"], sd["text_ln_final.bias"], cfg.ln_eps)
    return x @ sd["text_projection"]


# --------------------------------------------------------------------------------------------------
# sort head
# --------------------------------------------------------------------------------------------------
def sort_head(text, x, sd, cfg, prefix="pred_model."):
    """SortTransformer.forward, v2/model/sort_transformer.py:124-142 (+ blocks :35-80).
    text [B,nt,E] (detached transcripts), x [B,N,E] video tokens -> logits [B,nt,nt]."""
    E, h = x.shape[-1], cfg.sort_heads
    d = E // h
    N = x.shape[1]
    te = sd[prefix + "type_embed"]
    z = torch.cat([x + te[:, 0], text + te[:, 1]], 1)
    B, S, _ = z.shape
    for i in range(cfg.sort_depth):
        p = f"{prefix}blocks.{i}."
        y = layer_norm(z, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.sort_ln_eps)
        qkv = linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, S, 3, h, d)
        q = qkv[:, :, 0].transpose(1, 2) * (d ** -0.5)
        k = qkv[:, :, 1].transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        o = _softmax_av(q, k, v).transpose(1, 2).reshape(B, S, E)
        z = z + linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = layer_norm(z, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.sort_ln_eps)
        z = z + linear(gelu_erf(linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                       sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    y = layer_norm(z[:, N:], sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], cfg.sort_ln_eps)
    return linear(y, sd[prefix + "head.weight"], sd[prefix + "head.bias"])


# --------------------------------------------------------------------------------------------------
# model forward + losses
# --------------------------------------------------------------------------------------------------
def model_forward(sd, text, video, keep_ind, cfg):
    """TVTSv2_B_16.forward, model_dist_TVTSv2_ViT_B_16.py:61-95.
    text [n_trans*B, ctx] clip-major (row t*B+b).  Returns (text_emb [B,E], video_emb [B,E], pred_order|None)."""
    B = video.shape[0]
    t = text_tower(text, sd, cfg)                                  # [n_trans*B, E]
    t = t.reshape(-1, B, t.shape[-1])
    n_trans = t.shape[0]
    transcripts = t.detach().permute(1, 0, 2)                      # :69-70 no grad into text from the sort head
    text_emb = t.mean(0)                                           # :74-76
    if getattr(cfg, "post_mode", "all") == "h14":
        # TVTSv2_H_14.forward, model_dist_TVTSv2_ViT_H_14.py:97-131: the sort head sees the projected patch tokens only
        # (the reference runs this under fp16 autocast; the oracle restates the fp32 semantics)
        video_emb, vtok = video_tower(video, keep_ind, sd, cfg)
    else:
        vtok = video_tower(video, keep_ind, sd, cfg)
        video_emb = vtok[:, 0]
    pred = sort_head(transcripts, vtok, sd, cfg) if n_trans != 1 else None
    return text_emb, video_emb, pred


def sim_matrix(a, b, eps=1e-8):
    """model_dist_TVTSv2_ViT_B_16.py:119-127"""
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    return an @ bn.t()


def norm_softmax_loss(sim, temperature=0.05):
    """v2/model/loss.py:13-25"""
    z = sim / temperature
    i = torch.diagonal(torch.log_softmax(z, 1)).mean()
    j = torch.diagonal(torch.log_softmax(z.t(), 1)).mean()
    return -i - j


def sort_ce(pred, labels):
    """v2/trainer/trainer.py:487-492 : 2 * mean CE; labels int64 [B, n_trans]."""
    return 2.0 * F.cross_entropy(pred.reshape(-1, pred.shape[-1]), labels.reshape(-1).to(torch.long))


def step_losses(sd, text, video, keep_ind, labels, cfg, gather=None):
    """One trainer step's losses (v2/trainer/trainer.py:479-496).  `gather` (optional) maps the local
    [B,E] embeddings to the global [Bg,E] ones (AllGather_multi semantics, :41-57)."""
    te, ve, pred = model_forward(sd, text, video, keep_ind, cfg)
    if gather is not None:
        ve, te = gather(ve), gather(te)
    loss1 = norm_softmax_loss(sim_matrix(ve, te), cfg.temperature)
    loss2 = sort_ce(pred, labels) if pred is not None else torch.zeros((), dtype=loss1.dtype, device=loss1.device)
    return loss1, loss2, (te, ve, pred)


def step_with_grads(sd_tensors, text, video, keep_ind, labels, cfg, trainable=None):
    """fwd + bwd through autograd.  Returns (loss1, loss2, outputs, grads{name: tensor})."""
    sd = {}
    for k, v in sd_tensors.items():
        req = v.is_floating_point() and (trainable is None or k in trainable)
        sd[k] = v.detach().clone().requires_grad_(req)
    l1, l2, outs = step_losses(sd, text, video, keep_ind, labels, cfg)
    (l1 + l2).backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    return l1.detach(), l2.detach(), tuple(o.detach() if o is not None else None for o in outs), grads


# --------------------------------------------------------------------------------------------------
# TVTS v1 (BASELINE.json configs[4]): VideoMAE-style ViT with Conv3d tubelets + joint space-time attention, projection heads,
# the same sort head and losses.  The text encoder is HuggingFace DistilBERT (un-vendored dependency, `transformers` pin 4.10.2,
# parity unpinned): its [CLS] vectors are an INPUT of this restatement (`text_cls`), everything downstream is restated.
# --------------------------------------------------------------------------------------------------
def v1_video_tower(video, keep_ind, sd, cfg, prefix="video_model."):
    """VisionTransformer.forward_features, v1/model/video_encoder.py:178-217.
    video [B,T,3,H,W]; keep_ind [B, T/2, n] int64 (an independent subset per tube, v1/data_loader/YTTemporal_dataset.py:206-215).
    Conv3d(3, D, k=s=(2,p,p)) WITH bias (:89-91) on the [B,3,T,H,W] permutation == per-tubelet linear map over (c, dt, u, v)."""
    B, T, C, H, W = video.shape
    p, D = cfg.patch, cfg.width
    g = H // p
    nt = T // 2
    w = sd[prefix + "patch_embed.proj.weight"].reshape(D, -1)                # [D, 3*2*p*p], (c, dt, u, v) order
    x = video.reshape(B, nt, 2, C, g, p, g, p).permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(B, nt, g * g, C * 2 * p * p)
    keep = keep_ind[:, :nt].to(torch.long)
    idx = keep[..., None].expand(B, nt, keep.shape[-1], x.shape[-1])
    tok = torch.gather(x, 2, idx) @ w.t() + sd[prefix + "patch_embed.proj.bias"]          # gather before the (per-patch) linear map
    pos = sd[prefix + "pos_embed"][0]                                          # [P+1, D]
    tem = sd[prefix + "temporal_embed"][0]                                     # [num_tubes, D]
    tok = tok + pos[1:][keep] + tem[:nt][None, :, None, :]                     # (:193-199) added before the mask gather
    cls = (sd[prefix + "cls_token"][0, 0] + pos[0]).expand(B, 1, D)
    x = torch.cat([cls, tok.reshape(B, -1, D)], 1)
    h = cfg.heads
    d = D // h
    for i in range(cfg.layers):                                               # Block (:59-75): pre-LN, joint attention over all tokens
        q_ = f"{prefix}blocks.{i}."
        y = layer_norm(x, sd[q_ + "norm1.weight"], sd[q_ + "norm1.bias"], 1e-6)
        qkv = linear(y, sd[q_ + "attn.qkv.weight"], sd[q_ + "attn.qkv.bias"]).reshape(B, -1, 3, h, d)
        q = qkv[:, :, 0].transpose(1, 2) * (d ** -0.5)
        k = qkv[:, :, 1].transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        o = _softmax_av(q, k, v).transpose(1, 2).reshape(B, -1, D)
        x = x + linear(o, sd[q_ + "attn.proj.weight"], sd[q_ + "attn.proj.bias"])
        y = layer_norm(x, sd[q_ + "norm2.weight"], sd[q_ + "norm2.bias"], 1e-6)
        x = x + linear(gelu_erf(linear(y, sd[q_ + "mlp.fc1.weight"], sd[q_ + "mlp.fc1.bias"])), sd[q_ + "mlp.fc2.weight"], sd[q_ + "mlp.fc2.bias"])
    return layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-6)


def v1_model_forward(sd, text_cls, video, keep_ind, cfg):
    """TVTS.forward, v1/model/model_dist_TVTS.py:95-141, from the DistilBERT [CLS] vectors `text_cls` [n_trans*B, Dt] (clip-major).
    -> (text_emb [B,proj], video_emb [B,proj], pred_order [B,nt,nt] | None)."""
    B = video.shape[0]
    t = linear(torch.relu(text_cls), sd["txt_proj.1.weight"], sd["txt_proj.1.bias"])       # 'minimal' projection (:66-74)
    n_trans = text_cls.shape[0] // B
    transcripts = text_cls.detach().reshape(n_trans, B, -1).permute(1, 0, 2)
    text_emb = t.reshape(n_trans, B, -1).mean(0)
    vtok = v1_video_tower(video, keep_ind, sd, cfg)
    video_emb = linear(vtok[:, 0], sd["vid_proj.0.weight"], sd["vid_proj.0.bias"])
    pred = sort_head(transcripts, vtok, sd, cfg) if n_trans != 1 else None                 # raw 768-d tokens incl. CLS (:113-116)
    return text_emb, video_emb, pred


def distilbert_forward(sd, input_ids, attention_mask, heads, prefix="text_model."):
    """DistilBertModel.forward -> last_hidden_state [n, L, W]: the v1 text encoder (`AutoModel.from_pretrained('distilbert-base-uncased')`,
    v1/model/model_dist_TVTS.py:33; called at :124-126 with the tokenizer's input_ids / attention_mask).  UN-VENDORED DEPENDENCY
    (transformers==4.10.2 in v1's requirements; 5.5.0 installed here): restated from its published algorithm and pinned against the
    installed implementation by tests/test_oracle_golden.py (same weights, 1e-5).  Algorithm: embeddings = LayerNorm(word[ids] +
    position[0..L-1]) (eps 1e-12); per layer (POST-LN): q,k,v = three Linears; scores = (q / sqrt(d)) k^T with the padded KEY
    positions (attention_mask == 0) set to the most negative float; softmax; out_lin; x = sa_layer_norm(attn + x);
    x = output_layer_norm(lin2(gelu_erf(lin1(x))) + x).  Dropout p = 0.1 in the released config is active in training -- parity runs
    use p = 0 (as every fixture here does)."""
    n, L = input_ids.shape
    W = sd[prefix + "embeddings.word_embeddings.weight"].shape[1]
    d = W // heads
    x = sd[prefix + "embeddings.word_embeddings.weight"][input_ids.long()] + sd[prefix + "embeddings.position_embeddings.weight"][:L][None]
    x = layer_norm(x, sd[prefix + "embeddings.LayerNorm.weight"], sd[prefix + "embeddings.LayerNorm.bias"], 1e-12)
    key_ok = attention_mask.bool()[:, None, None, :]                                   # [n, 1, 1, L]
    i = 0
    while f"{prefix}transformer.layer.{i}.attention.q_lin.weight" in sd:
        p_ = f"{prefix}transformer.layer.{i}."
        q = linear(x, sd[p_ + "attention.q_lin.weight"], sd[p_ + "attention.q_lin.bias"]).reshape(n, L, heads, d).transpose(1, 2)
        k = linear(x, sd[p_ + "attention.k_lin.weight"], sd[p_ + "attention.k_lin.bias"]).reshape(n, L, heads, d).transpose(1, 2)
        v = linear(x, sd[p_ + "attention.v_lin.weight"], sd[p_ + "attention.v_lin.bias"]).reshape(n, L, heads, d).transpose(1, 2)
        s_ = (q * (d ** -0.5)) @ k.transpose(-1, -2)
        s_ = s_.masked_fill(~key_ok, torch.finfo(s_.dtype).min)
        o = (torch.softmax(s_, -1) @ v).transpose(1, 2).reshape(n, L, W)
        x = layer_norm(linear(o, sd[p_ + "attention.out_lin.weight"], sd[p_ + "attention.out_lin.bias"]) + x,
                       sd[p_ + "sa_layer_norm.weight"], sd[p_ + "sa_layer_norm.bias"], 1e-12)
        f_ = linear(gelu_erf(linear(x, sd[p_ + "ffn.lin1.weight"], sd[p_ + "ffn.lin1.bias"])), sd[p_ + "ffn.lin2.weight"], sd[p_ + "ffn.lin2.bias"])
        x = layer_norm(f_ + x, sd[p_ + "output_layer_norm.weight"], sd[p_ + "output_layer_norm.bias"], 1e-12)
        i += 1
    return x


def v1_step_with_grads(sd_tensors, text, video, keep_ind, labels, cfg, text_heads):
    """One v1 training step (v1/trainer/trainer.py:135-155 at world size 1) INCLUDING the DistilBERT text encoder:
    text = {'input_ids', 'attention_mask'} [n_trans*B, L] clip-major.  -> loss1, loss2, (text_emb, video_emb, pred), {name: grad}."""
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd_tensors.items()}
    hidden = distilbert_forward(sd, text["input_ids"], text["attention_mask"], text_heads)
    te, ve, pred = v1_model_forward(sd, hidden[:, 0, :], video, keep_ind, cfg)
    loss1 = norm_softmax_loss(sim_matrix(ve, te), 0.05)
    loss2 = sort_ce(pred, labels) if pred is not None else torch.zeros(())
    (loss1 + loss2).backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    return loss1.detach(), loss2.detach(), (te.detach(), ve.detach(), None if pred is None else pred.detach()), grads


# --------------------------------------------------------------------------------------------------
# optimiser: transformers==4.10.2 AdamW (un-vendored; parity unpinned by the reference -- restated from its
# published algorithm: Adam with bias correction, eps added to sqrt(v) BEFORE bias correction is folded
# into the step size, decoupled weight decay applied AFTER the Adam update with the un-corrected lr).
# Call site: v2/train_dist_TVTSv2_ViT_B_16.py:119-125.
# --------------------------------------------------------------------------------------------------
def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
