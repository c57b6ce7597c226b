"""Deterministic synthetic weights and batches.

There is no network for checkpoints or datasets, so every parity test and the benchmark use
 * a seeded random state_dict carrying the reference's parameter names/shapes (CLIP-style stds;
   `timeattn.*` re-drawn N(0, 0.02^2) instead of the reference's zero init so the temporal branch is
   exercised -- v2/model/video_encoder_ViT_B_16.py:28-34 makes it exactly 0 at step 0), and
 * seeded batches shaped like the trainer's: video ~ N(0,1), a random tube-mask `keep_ind`
   (v2/data_loader/YTTemporal_dataset.py:207-213), CLIP-like token rows [SOT, r_1..r_l, EOT, 0...]
   in clip-major order (v2/trainer/trainer.py:465-473), labels = arange(n_trans) (:149 of the dataset).
The CPU generators are bit-reproducible across machines, so goldens made in the build container
match inputs regenerated on the GPU box.
"""
import numpy as np
import torch


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def _uniform(g, *shape, bound=1.0):
    return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def make_state_dict(cfg, seed=1234, zero_timeattn=False):
    """Reference-named fp32 state_dict for TVTSv2_{B_16,B_32} (and tiny variants)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    D, E, W = cfg.width, cfg.embed_dim, cfg.text_width
    P = cfg.patches_per_frame

    def ln(name, dim):
        sd[name + ".weight"] = 1.0 + _randn(g, dim, std=0.1)
        sd[name + ".bias"] = _randn(g, dim, std=0.02)

    # ---- text tower (names from model_dist_TVTSv2_ViT_B_16.py:22-26) ----
    sd["text_positional_embedding"] = _randn(g, cfg.context, W, std=0.01)
    sd["text_projection"] = _randn(g, W, E, std=W ** -0.5)
    proj_std = (W ** -0.5) * ((2 * cfg.text_layers) ** -0.5)
    for i in range(cfg.text_layers):
        p = f"text_model.resblocks.{i}."
        sd[p + "attn.in_proj_weight"] = _randn(g, 3 * W, W, std=W ** -0.5)
        sd[p + "attn.in_proj_bias"] = _randn(g, 3 * W, std=0.02)
        sd[p + "attn.out_proj.weight"] = _randn(g, W, W, std=proj_std)
        sd[p + "attn.out_proj.bias"] = _randn(g, W, std=0.02)
        ln(p + "ln_1", W)
        sd[p + "mlp.c_fc.weight"] = _randn(g, 4 * W, W, std=(2 * W) ** -0.5)
        sd[p + "mlp.c_fc.bias"] = _randn(g, 4 * W, std=0.02)
        sd[p + "mlp.c_proj.weight"] = _randn(g, W, 4 * W, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = _randn(g, W, std=0.02)
        ln(p + "ln_2", W)
    sd["text_token_embedding.weight"] = _randn(g, cfg.vocab, W, std=0.02)
    ln("text_ln_final", W)

    # ---- video tower (video_encoder_ViT_B_16.py:147-174) ----
    v = "video_model."
    sc = D ** -0.5
    sd[v + "class_embedding"] = _randn(g, D, std=sc)
    sd[v + "positional_embedding"] = _randn(g, P + 1, D, std=sc)
    sd[v + "proj"] = _randn(g, D, E, std=sc)
    sd[v + "temporal_embedding"] = _randn(g, cfg.num_frames, D, std=sc)
    sd[v + "conv1.weight"] = _randn(g, D, 3, cfg.patch, cfg.patch, std=(3 * cfg.patch ** 2) ** -0.5)
    ln(v + "ln_pre", D)
    vproj_std = (D ** -0.5) * ((2 * cfg.layers) ** -0.5)
    for i in range(cfg.layers):
        p = f"{v}transformer.resblocks.{i}."
        for a in ("attn", "timeattn"):
            z = zero_timeattn and a == "timeattn"
            sd[p + a + ".qkv.weight"] = torch.zeros(3 * D, D) if z else _randn(g, 3 * D, D, std=(D ** -0.5 if a == "attn" else 0.02))
            sd[p + a + ".qkv.bias"] = torch.zeros(3 * D) if z else _randn(g, 3 * D, std=0.02)
            sd[p + a + ".proj.weight"] = torch.ones(D, D) if z else _randn(g, D, D, std=(vproj_std if a == "attn" else 0.02))
            sd[p + a + ".proj.bias"] = torch.zeros(D) if z else _randn(g, D, std=0.02)
        ln(p + "ln_3", D)
        ln(p + "ln_1", D)
        sd[p + "mlp.c_fc.weight"] = _randn(g, 4 * D, D, std=(2 * D) ** -0.5)
        sd[p + "mlp.c_fc.bias"] = _randn(g, 4 * D, std=0.02)
        sd[p + "mlp.c_proj.weight"] = _randn(g, D, 4 * D, std=vproj_std)
        sd[p + "mlp.c_proj.bias"] = _randn(g, D, std=0.02)
        ln(p + "ln_2", D)
    ln(v + "ln_post", D)

    # ---- sort head (sort_transformer.py:82-113; default nn.Linear init since _init_weights is never applied) ----
    s = "pred_model."
    sd[s + "type_embed"] = _randn(g, 1, 2, E, std=0.02)
    for i in range(cfg.sort_depth):
        p = f"{s}blocks.{i}."
        ln(p + "norm1", E)
        sd[p + "attn.qkv.weight"] = _uniform(g, 3 * E, E, bound=E ** -0.5)
        sd[p + "attn.qkv.bias"] = _uniform(g, 3 * E, bound=E ** -0.5)
        sd[p + "attn.proj.weight"] = _uniform(g, E, E, bound=E ** -0.5)
        sd[p + "attn.proj.bias"] = _uniform(g, E, bound=E ** -0.5)
        ln(p + "norm2", E)
        sd[p + "mlp.fc1.weight"] = _uniform(g, 4 * E, E, bound=E ** -0.5)
        sd[p + "mlp.fc1.bias"] = _uniform(g, 4 * E, bound=E ** -0.5)
        sd[p + "mlp.fc2.weight"] = _uniform(g, E, 4 * E, bound=(4 * E) ** -0.5)
        sd[p + "mlp.fc2.bias"] = _uniform(g, E, bound=(4 * E) ** -0.5)
    ln(s + "norm", E)
    sd[s + "head.weight"] = _uniform(g, cfg.n_trans, E, bound=E ** -0.5)
    sd[s + "head.bias"] = _uniform(g, cfg.n_trans, bound=E ** -0.5)
    return sd


def make_keep_ind(cfg, batch, seed=0):
    """[B, n] int64: per-sample random subset (unsorted permutation prefix), shared by all frames."""
    P, n = cfg.patches_per_frame, cfg.kept_per_frame
    rows = [np.random.RandomState(seed + b).permutation(P)[:n] for b in range(batch)]
    return torch.from_numpy(np.stack(rows).astype(np.int64))


def make_tokens(cfg, n_rows, seed=0, dtype=torch.int32):
    """[n_rows, ctx] CLIP-style rows: SOT, l random ids, EOT (= max id), zero padding."""
    rs = np.random.RandomState(seed + 7919)
    sot, eot = cfg.vocab - 2, cfg.vocab - 1
    out = np.zeros((n_rows, cfg.context), dtype=np.int64)
    hi = min(40, cfg.context - 2)
    for r in range(n_rows):
        l = int(rs.randint(6, hi + 1))
        out[r, 0] = sot
        out[r, 1:1 + l] = rs.randint(1, sot, size=l)
        out[r, 1 + l] = eot
    return torch.from_numpy(out).to(dtype)


def make_batch(cfg, batch, frames, n_trans=4, seed=0, rank=0):
    """dict like the trainer's `data` after tokenisation (v2/trainer/trainer.py:463-475)."""
    s = seed + 1000003 * rank
    g = torch.Generator().manual_seed(s)
    video = torch.randn(batch, frames, 3, cfg.resolution, cfg.resolution, generator=g, dtype=torch.float32)
    return {
        "video": video,
        "keep_ind": make_keep_ind(cfg, batch, seed=s),
        "text": make_tokens(cfg, n_trans * batch, seed=s),
        "label": torch.arange(n_trans, dtype=torch.int64).repeat(batch, 1),
    }
