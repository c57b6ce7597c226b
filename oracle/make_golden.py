"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python oracle/make_golden.py            # writes tests/golden/{c1_b32,tiny_B,tiny_B_mask,tiny_B_cap,tiny_H,tiny_H640,tiny_ds,tiny_v1,tiny_v1_full}.npz

For each case: seeded state_dict (tvts_b200.synthetic.make_state_dict) is loaded strict=True into the
reference modules, the reference forward + the trainer's loss lines (v2/trainer/trainer.py:479-496)
run on a seeded batch, and outputs + per-parameter gradient fingerprints are stored.  The fixtures
are small (no weights: they are regenerated from the seed on the GPU box).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_shims  # noqa: E402
from tvts_b200 import config as C  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402


def build_reference_model(cfg):
    """A reference TVTSv2_B_* instance with cfg's dims (the class hard-codes B/16|B/32 dims in __init__,
    so for tiny dims the submodules are built with the reference constructors and attached to an
    un-initialised instance; forward()/compute_text()/compute_video() are the reference's own)."""
    from torch import nn
    import model.model_dist_TVTSv2_ViT_B_32 as ref_model
    from model.video_encoder_ViT_B_32 import VisionTransformer
    from model.sort_transformer import SortTransformer
    from CLIP.clip.model import CLIP

    m = ref_model.TVTSv2_B_32.__new__(ref_model.TVTSv2_B_32)
    nn.Module.__init__(m)
    clip_model = CLIP(cfg.embed_dim, cfg.resolution, 1, 64, cfg.patch, cfg.context, cfg.vocab,
                      cfg.text_width, cfg.text_heads, cfg.text_layers)
    m.text_model = clip_model.transformer
    m.text_token_embedding = clip_model.token_embedding
    m.text_positional_embedding = clip_model.positional_embedding
    m.text_ln_final = clip_model.ln_final
    m.text_projection = clip_model.text_projection
    m.video_model = VisionTransformer(input_resolution=cfg.resolution, patch_size=cfg.patch, width=cfg.width,
                                      layers=cfg.layers, heads=cfg.heads, output_dim=cfg.embed_dim,
                                      num_frames=cfg.num_frames, mask_ratio=cfg.mask_ratio)
    m.n_trans = cfg.n_trans
    m.pred_model = SortTransformer(num_classes=cfg.n_trans, embed_dim=cfg.embed_dim, num_heads=cfg.sort_heads)
    return m, ref_model.sim_matrix


def build_reference_model_h14(cfg):
    """A reference TVTSv2_H_14 instance with cfg's (tiny) dims: the modified-OpenCLIP video encoder
    (v2/model/video_encoder_ViT_H_14.py), the OpenCLIP text Transformer with its causal attn_mask
    (v2/OpenCLIP/transformer.py), the SortTransformer; forward()/compute_text()/compute_video() are the reference's own."""
    from torch import nn
    import model.model_dist_TVTSv2_ViT_H_14 as ref_model
    from model.video_encoder_ViT_H_14 import VisionTransformer, LayerNorm
    from model.sort_transformer import SortTransformer
    from OpenCLIP.transformer import Transformer as TextTransformer, LayerNorm as TextLayerNorm

    m = ref_model.TVTSv2_H_14.__new__(ref_model.TVTSv2_H_14)
    nn.Module.__init__(m)
    W = cfg.text_width
    m.text_model = TextTransformer(width=W, layers=cfg.text_layers, heads=cfg.text_heads, act_layer=nn.GELU)
    m.text_token_embedding = nn.Embedding(cfg.vocab, W)
    m.text_positional_embedding = nn.Parameter(torch.empty(cfg.context, W))
    m.text_ln_final = TextLayerNorm(W)
    m.text_projection = nn.Parameter(torch.empty(W, cfg.embed_dim))
    mask = torch.empty(cfg.context, cfg.context)
    mask.fill_(float("-inf"))
    mask.triu_(1)                                               # OpenCLIP CLIP.build_attention_mask (v2/OpenCLIP/model.py)
    m.text_attn_mask = mask
    m.video_model = VisionTransformer(image_size=cfg.resolution, patch_size=cfg.patch, width=cfg.width, layers=cfg.layers,
                                      heads=cfg.heads, mlp_ratio=4.0, output_dim=cfg.embed_dim, act_layer=nn.GELU,
                                      norm_layer=LayerNorm, num_frames=cfg.num_frames, mask_ratio=cfg.mask_ratio)
    m.n_trans = cfg.n_trans
    m.pred_model = SortTransformer(num_classes=cfg.n_trans, embed_dim=cfg.embed_dim, num_heads=cfg.sort_heads)
    return m, ref_model.sim_matrix


def run_case(name, cfg, batch, frames, n_trans, seed):
    from model.loss import NormSoftmaxLoss
    m, sim_matrix = build_reference_model_h14(cfg) if cfg.post_mode == "h14" else build_reference_model(cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)
    m.train()
    data = make_batch(cfg, batch, frames, n_trans=n_trans, seed=seed)
    text_e, video_e, pred = m(data)
    # trainer lines 481-494 at world_size 1 (all_gather of one rank is the identity)
    loss1 = NormSoftmaxLoss(cfg.temperature)(sim_matrix(video_e, text_e))
    if pred is not None:
        loss2 = torch.nn.CrossEntropyLoss()(pred.reshape(-1, pred.shape[-1]), data["label"].reshape(-1)) * 2
    else:
        loss2 = torch.zeros(())
    (loss1 + loss2).backward()
    out = {
        "text_emb": text_e.detach().numpy(), "video_emb": video_e.detach().numpy(),
        "loss1": np.float64(loss1.item()), "loss2": np.float64(loss2.item()),
        "batch": batch, "frames": frames, "n_trans": n_trans, "seed": seed,
    }
    if pred is not None:
        out["pred_order"] = pred.detach().numpy()
    names, norms, heads = [], [], []
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        heads.append(p.grad.reshape(-1)[:8].double().numpy().copy() if p.numel() >= 8
                     else np.pad(p.grad.reshape(-1).double().numpy(), (0, 8 - p.numel())))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    out["grad_heads"] = np.stack(heads)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)

    # immediately pin the restatement
    import tvts_oracle as O
    l1, l2, (te, ve, pr), grads = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    err = lambda a, b: float((a - b).abs().max())
    print(f"[{name}] loss1 ref {loss1.item():.7f} oracle {l1.item():.7f} | loss2 ref {loss2.item():.7f} oracle {l2.item():.7f}")
    print(f"   max|d text_emb| {err(te, text_e.detach()):.2e}  max|d video_emb| {err(ve, video_e.detach()):.2e}"
          + (f"  max|d pred| {err(pr, pred.detach()):.2e}" if pred is not None else ""))
    worst = 0.0
    for k, nrm in zip(names, norms):
        gn = grads[k].double().norm().item() if k in grads else 0.0
        worst = max(worst, abs(gn - nrm) / (nrm + 1e-12))
    print(f"   worst relative grad-norm deviation over {len(names)} params: {worst:.2e}   -> {path} ({os.path.getsize(path)} B)")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_shims.install("v2")
    run_case("tiny_B", C.TINY_B, batch=3, frames=2, n_trans=4, seed=11)
    run_case("tiny_B_mask", C.TINY_B_MASK, batch=2, frames=3, n_trans=4, seed=12)
    run_case("tiny_B_cap", C.TINY_B, batch=4, frames=2, n_trans=1, seed=13)       # caption mode: pred_order None
    run_case("c1_b32", C.TVTSV2_B_32, batch=4, frames=2, n_trans=4, seed=0)       # BASELINE.json configs[0]
    run_case("tiny_H", C.TINY_H, batch=2, frames=3, n_trans=4, seed=14)           # H/14 semantics (configs[3]) at toy dims
    run_case("tiny_H640", C.TINY_H640, batch=2, frames=3, n_trans=4, seed=17)     # ... at the smallest width the CUDA kernels take
    run_case_downstream()                                                         # v2/downstream towers (mask 0, no sort head), forward only
    run_case_v1()                                                                 # TVTS v1 semantics (configs[4]) at toy dims
    run_case_v1_full()                                                            # ... including the DistilBERT text encoder, head dim 64


# ------------------------------------------------------------------------------------------------------------------ TVTS v1
from make_golden_spec import spec_state_dict  # noqa: E402


def run_case_v1(name="tiny_v1", seed=15):
    """TVTS v1 at toy dims through the UNMODIFIED reference classes (v1/model/model_dist_TVTS.py, video_encoder.py): a randomly
    initialised 1-layer DistilBERT from the installed `transformers` stands in for 'distilbert-base-uncased'; its [CLS] vectors are
    stored in the fixture because the oracle takes them as input (the text encoder is an un-vendored dependency)."""
    from functools import partial
    from torch import nn
    import transformers
    ref_shims.install("v1")
    import model.model_dist_TVTS as ref_model
    from model.video_encoder import VisionTransformer
    from model.sort_transformer import SortTransformer
    from model.loss import NormSoftmaxLoss

    D, heads, depth, patch, res, frames, nt, B, proj = 96, 2, 2, 16, 64, 8, 4, 2, 32
    torch.manual_seed(seed)
    m = ref_model.TVTS.__new__(ref_model.TVTS)
    nn.Module.__init__(m)
    m.text_params = {"model": "distilbert-base-uncased", "pretrained": True}
    m.text_model = transformers.DistilBertModel(transformers.DistilBertConfig(vocab_size=128, dim=D, n_layers=1, n_heads=2, hidden_dim=2 * D,
                                                                              max_position_embeddings=32, dropout=0.0, attention_dropout=0.0))
    m.video_model = VisionTransformer(img_size=res, patch_size=patch, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), num_frames=frames)
    m.video_model.pre_logits = nn.Identity()
    m.txt_proj = nn.Sequential(nn.ReLU(), nn.Linear(D, proj))
    m.vid_proj = nn.Sequential(nn.Linear(D, proj))
    m.n_trans = nt
    m.pred_model = SortTransformer(num_classes=nt, embed_dim=D, num_heads=heads)
    names = [k for k, _ in m.named_parameters() if not k.startswith("text_model.")]
    shapes = [tuple(p.shape) for k, p in m.named_parameters() if not k.startswith("text_model.")]
    sd = spec_state_dict(names, shapes, 4321)
    m.load_state_dict(sd, strict=False)
    m.train()
    g = torch.Generator().manual_seed(seed)
    P = (res // patch) ** 2
    n_keep = P // 2
    video = torch.randn(B, frames, 3, res, res, generator=torch.Generator().manual_seed(seed + 1))   # regenerated from the seed by the test
    keep = torch.stack([torch.stack([torch.randperm(P, generator=g)[:n_keep] for _ in range(frames // 2)]) for _ in range(B)])
    ids = torch.randint(1, 128, (nt * B, 12), generator=g)
    text = {"input_ids": ids, "attention_mask": torch.ones_like(ids)}
    labels = torch.arange(nt).repeat(B, 1)
    text_cls = m.text_model(**text).last_hidden_state[:, 0, :].detach()
    te, ve, pred = m({"text": text, "video": video, "keep_ind": keep})
    loss1 = NormSoftmaxLoss(0.05)(ref_model.sim_matrix(ve, te))
    loss2 = torch.nn.CrossEntropyLoss()(pred.reshape(-1, nt), labels.reshape(-1)) * 2
    (loss1 + loss2).backward()
    gn = {k: p.grad.double().norm().item() for k, p in m.named_parameters() if p.grad is not None and not k.startswith("text_model.")}
    out = dict(names=np.array(names), shapes=np.array([",".join(map(str, s)) for s in shapes]), wseed=4321, text_cls=text_cls.numpy(),
               video_seed=seed + 1, keep_ind=keep.numpy(), text_emb=te.detach().numpy(), video_emb=ve.detach().numpy(),
               pred_order=pred.detach().numpy(), loss1=np.float64(loss1.item()), loss2=np.float64(loss2.item()),
               grad_names=np.array(list(gn)), grad_norms=np.array(list(gn.values())),
               dims=np.array([D, heads, depth, patch, res, frames, nt, B, proj]))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    import tvts_oracle as O
    cfg = types.SimpleNamespace(patch=patch, width=D, heads=heads, layers=depth, sort_heads=heads, sort_depth=2, sort_ln_eps=1e-6)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ote, ove, opr = O.v1_model_forward(sdr, text_cls, video, keep, cfg)
    l1 = O.norm_softmax_loss(O.sim_matrix(ove, ote), 0.05)
    l2 = O.sort_ce(opr, labels)
    (l1 + l2).backward()
    worst = max(abs(sdr[k].grad.double().norm().item() - v) / (v + 1e-12) for k, v in gn.items() if k != "txt_proj.1.weight" or True)
    print(f"[{name}] loss1 ref {loss1.item():.7f} oracle {l1.item():.7f} | loss2 ref {loss2.item():.7f} oracle {l2.item():.7f}")
    print(f"   max|d text_emb| {(ote - te).abs().max().item():.2e} max|d video_emb| {(ove - ve).abs().max().item():.2e} "
          f"max|d pred| {(opr - pred).abs().max().item():.2e}  worst rel grad-norm dev over {len(gn)} params {worst:.2e} -> {path} ({os.path.getsize(path)} B)")


# ------------------------------------------------------------------------------------------------------------------ downstream
def run_case_downstream(name="tiny_ds", seed=16):
    """v2/downstream/model_TVTSv2_ViT_B_32.py (+ _mc) at toy dims through the UNMODIFIED reference classes: the towers of TINY_B with
    mask_ratio 0 and no sort head, forward only (eval, no_grad), n_trans 3.  Weights = the pre-training seeded state_dict minus
    pred_model.* (what a released downstream checkpoint holds)."""
    from torch import nn
    ref_shims.install("v2")
    import downstream.model_TVTSv2_ViT_B_32 as ds
    import downstream.model_TVTSv2_ViT_B_32_mc as ds_mc
    from model.video_encoder_ViT_B_32 import VisionTransformer
    from CLIP.clip.model import CLIP

    cfg = C.TINY_B
    sd = {k: v for k, v in make_state_dict(cfg, seed=1234).items() if not k.startswith("pred_model.")}
    B, T, nt = 3, 2, 3
    data = make_batch(cfg, B, T, n_trans=nt, seed=seed)
    out = {"batch": B, "frames": T, "n_trans": nt, "seed": seed}
    for tag, mod in (("", ds), ("_mc", ds_mc)):
        m = mod.TVTSv2_B_32.__new__(mod.TVTSv2_B_32)
        nn.Module.__init__(m)
        clip_model = CLIP(cfg.embed_dim, cfg.resolution, 1, 64, cfg.patch, cfg.context, cfg.vocab, cfg.text_width, cfg.text_heads, cfg.text_layers)
        m.text_model = clip_model.transformer
        m.text_token_embedding = clip_model.token_embedding
        m.text_positional_embedding = clip_model.positional_embedding
        m.text_ln_final = clip_model.ln_final
        m.text_projection = clip_model.text_projection
        m.video_model = VisionTransformer(input_resolution=cfg.resolution, patch_size=cfg.patch, width=cfg.width, layers=cfg.layers,
                                          heads=cfg.heads, output_dim=cfg.embed_dim, num_frames=cfg.num_frames, mask_ratio=0.)
        m.load_state_dict(sd, strict=True)
        m.eval()
        with torch.no_grad():
            te, ve = m(data, return_embeds=True)
            out["text_emb" + tag], out["video_emb" + tag] = te.numpy(), ve.numpy()
            if tag == "":
                out["sims"] = m(data, return_embeds=False).numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    import tvts_oracle as O
    full = make_state_dict(cfg, seed=1234)
    ote, ove, _ = O.model_forward(full, data["text"], data["video"], data["keep_ind"], cfg)
    print(f"[{name}] max|d text_emb| {(ote - torch.from_numpy(out['text_emb'])).abs().max().item():.2e} "
          f"max|d video_emb| {(ove - torch.from_numpy(out['video_emb'])).abs().max().item():.2e}  text_emb_mc {out['text_emb_mc'].shape} -> {path} ({os.path.getsize(path)} B)")


def run_case_v1_full(name="tiny_v1_full", seed=18):
    """TVTS v1 INCLUDING the text encoder, at toy dims whose head dim is 64 (what the CUDA attention kernels take): video ViT 128 wide /
    2 heads / 2 blocks on 64x64 frames (16 patches, 8 kept per tube), a randomly initialised 2-layer DistilBERT (dim 128, 2 heads, dropout 0)
    from the installed `transformers` in place of 'distilbert-base-uncased', right-padded captions.  Every parameter (text encoder
    included) comes from spec_state_dict(names, shapes, seed), so the fixture only stores names / shapes / results."""
    from functools import partial
    from torch import nn
    import transformers
    ref_shims.install("v1")
    import model.model_dist_TVTS as ref_model
    from model.video_encoder import VisionTransformer
    from model.sort_transformer import SortTransformer
    from model.loss import NormSoftmaxLoss

    D, heads, depth, patch, res, frames, nt, B, proj, Lc, vocab = 128, 2, 2, 16, 64, 8, 4, 2, 64, 12, 128
    torch.manual_seed(seed)
    m = ref_model.TVTS.__new__(ref_model.TVTS)
    nn.Module.__init__(m)
    m.text_params = {"model": "distilbert-base-uncased", "pretrained": True}
    m.text_model = transformers.DistilBertModel(transformers.DistilBertConfig(vocab_size=vocab, dim=D, n_layers=2, n_heads=heads, hidden_dim=2 * D,
                                                                              max_position_embeddings=32, dropout=0.0, attention_dropout=0.0))
    m.video_model = VisionTransformer(img_size=res, patch_size=patch, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), num_frames=frames)
    m.video_model.pre_logits = nn.Identity()
    m.txt_proj = nn.Sequential(nn.ReLU(), nn.Linear(D, proj))
    m.vid_proj = nn.Sequential(nn.Linear(D, proj))
    m.n_trans = nt
    m.pred_model = SortTransformer(num_classes=nt, embed_dim=D, num_heads=heads)
    names = [k for k, _ in m.named_parameters()]
    shapes = [tuple(p.shape) for _, p in m.named_parameters()]
    sd = spec_state_dict(names, shapes, 4321)
    m.load_state_dict(sd, strict=True)
    m.train()
    g = torch.Generator().manual_seed(seed)
    P = (res // patch) ** 2
    n_keep = P // 2
    video = torch.randn(B, frames, 3, res, res, generator=torch.Generator().manual_seed(seed + 1))
    keep = torch.stack([torch.stack([torch.randperm(P, generator=g)[:n_keep] for _ in range(frames // 2)]) for _ in range(B)])
    ids = torch.randint(1, vocab, (nt * B, Lc), generator=g)
    lens = torch.randint(3, Lc + 1, (nt * B,), generator=g)
    lens[0] = Lc
    mask = (torch.arange(Lc)[None, :] < lens[:, None]).long()
    ids = ids * mask                                                         # [PAD] = 0 on the right, like the tokenizer
    text = {"input_ids": ids, "attention_mask": mask}
    labels = torch.arange(nt).repeat(B, 1)
    te, ve, pred = m({"text": text, "video": video, "keep_ind": keep})
    loss1 = NormSoftmaxLoss(0.05)(ref_model.sim_matrix(ve, te))
    loss2 = torch.nn.CrossEntropyLoss()(pred.reshape(-1, nt), labels.reshape(-1)) * 2
    (loss1 + loss2).backward()
    gn = {k: p.grad.double().norm().item() for k, p in m.named_parameters() if p.grad is not None}
    out = dict(names=np.array(names), shapes=np.array([",".join(map(str, s)) for s in shapes]), wseed=4321, video_seed=seed + 1,
               keep_ind=keep.numpy(), input_ids=ids.numpy(), attention_mask=mask.numpy(), text_emb=te.detach().numpy(),
               video_emb=ve.detach().numpy(), pred_order=pred.detach().numpy(), loss1=np.float64(loss1.item()), loss2=np.float64(loss2.item()),
               grad_names=np.array(list(gn)), grad_norms=np.array(list(gn.values())),
               dims=np.array([D, heads, depth, patch, res, frames, nt, B, proj, Lc, vocab]))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    import tvts_oracle as O
    cfg = types.SimpleNamespace(patch=patch, width=D, heads=heads, layers=depth, sort_heads=heads, sort_depth=2, sort_ln_eps=1e-6)
    l1, l2, (ote, ove, opr), grads = O.v1_step_with_grads(sd, text, video, keep, labels, cfg, heads)
    # (k_lin.bias has an exactly-zero true gradient -- softmax is invariant to a constant added to every key -- so only noise ~1e-9 there)
    worst = max(abs(grads[k].double().norm().item() - v) / (v + 1e-12) for k, v in gn.items() if v > 1e-6)
    print(f"[{name}] loss1 ref {loss1.item():.7f} oracle {l1.item():.7f} | loss2 ref {loss2.item():.7f} oracle {l2.item():.7f}")
    print(f"   max|d text_emb| {(ote - te).abs().max().item():.2e} max|d video_emb| {(ove - ve).abs().max().item():.2e} "
          f"max|d pred| {(opr - pred).abs().max().item():.2e}  worst rel grad-norm dev over {len(gn)} params (norm > 1e-6) {worst:.2e} -> {path} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    if "--downstream" in sys.argv:      # regenerate only tests/golden/tiny_ds.npz
        torch.manual_seed(0)
        run_case_downstream()
    elif "--v1full" in sys.argv:        # regenerate only tests/golden/tiny_v1_full.npz
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        run_case_v1_full()
    elif "--h640" in sys.argv:          # regenerate only tests/golden/tiny_H640.npz
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        ref_shims.install("v2")
        run_case("tiny_H640", C.TINY_H640, batch=2, frames=3, n_trans=4, seed=17)
    else:
        main()
