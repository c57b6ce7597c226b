"""The drop-in package tree (tvts_b200/dropin) exposes the reference's module paths / class names for the hot path, and the
model keeps the reference's parameter names (v2/train_dist_TVTSv2_ViT_B_16.py:8-20,66-107 rely on both)."""
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve():
    code = ("import model.model_dist_TVTSv2_ViT_B_16 as a, model.model_dist_TVTSv2_ViT_B_32 as b, model.loss as l, "
            "trainer.trainer as t, model.sort_transformer as s, model.video_encoder_ViT_B_16 as v;"
            "assert a.TVTSv2_B_16 and b.TVTSv2_B_32 and a.sim_matrix and l.NormSoftmaxLoss and t.AllGather_multi and s.SortTransformer and v.VisionTransformer;"
            # everything else the entry / downstream scripts import from the shadowed packages (train_dist_TVTSv2_ViT_*.py:8-20, zero_ret_*:9,17)
            "import model.metric as mm, model.model_dist_TVTSv2_ViT_H_14 as h, model.video_encoder_ViT_H_14 as vh, model.video_encoder_ViT_B_32 as v32;"
            "from trainer import Trainer_TVTSv2_B_16, Trainer_TVTSv2_B_32, Trainer_TVTSv2_H_14;"
            "from trainer.trainer import verbose, format_nested_metrics_for_writer;"
            "assert mm.t2v_metrics and mm.v2t_metrics and h.TVTSv2_H_14 and vh.VisionTransformer and vh.LayerNorm and vh.QuickGELU and v32.VisionTransformer")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tvts_b200", "dropin") + os.pathsep + ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")


def test_parameter_names_match_reference_layout():
    from tvts_b200 import config as C
    from tvts_b200 import modules as M
    from tvts_b200.synthetic import make_state_dict
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=C.TINY_B)
    names = set(dict(m.named_parameters()).keys())
    assert names == set(make_state_dict(C.TINY_B).keys())
    for k in ("video_model.transformer.resblocks.0.timeattn.qkv.weight", "video_model.transformer.resblocks.1.ln_3.bias",
              "video_model.conv1.weight", "video_model.temporal_embedding", "video_model.proj",
              "text_model.resblocks.0.attn.in_proj_weight", "text_model.resblocks.1.attn.out_proj.bias", "text_token_embedding.weight",
              "text_positional_embedding", "text_ln_final.weight", "text_projection", "pred_model.type_embed",
              "pred_model.blocks.1.mlp.fc2.weight", "pred_model.norm.bias", "pred_model.head.weight"):
        assert k in names, k


def test_downstream_import_paths_resolve():
    code = ("import downstream.model_TVTSv2_ViT_B_16 as a, downstream.model_TVTSv2_ViT_B_32 as b, "
            "downstream.model_TVTSv2_ViT_B_16_mc as c, downstream.model_TVTSv2_ViT_B_32_mc as d;"
            "assert a.TVTSv2_B_16 and b.TVTSv2_B_32 and c.TVTSv2_B_16 and d.TVTSv2_B_32 and a.sim_matrix and d.sim_matrix;"
            "assert a.TVTSv2_B_16.MEAN_OVER_CLIPS and not c.TVTSv2_B_16.MEAN_OVER_CLIPS")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tvts_b200", "dropin") + os.pathsep + ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")


def test_downstream_model_matches_reference_golden(emu_backend):
    """v2/downstream/model_TVTSv2_ViT_B_32.py (+_mc): no sort head / no pred_model.* keys, mask_ratio 0, (text, video) return,
    sim_matrix when return_embeds=False -- against the fixture written by the executed reference (oracle/make_golden.py)."""
    import numpy as np
    import torch
    from tvts_b200 import config as C
    from tvts_b200 import engine as E
    from tvts_b200 import modules as M
    from tvts_b200.synthetic import make_batch, make_state_dict
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_ds.npz"))
    cfg = C.TINY_B
    sd = {k: v for k, v in make_state_dict(cfg, seed=1234).items() if not k.startswith("pred_model.")}
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    for cls, tag in ((M.TVTSv2_B_32_downstream, ""), (M.TVTSv2_B_32_downstream_mc, "_mc")):
        E.WEIGHTS.clear()
        m = cls(arch=cfg)
        assert not hasattr(m, "pred_model") and m.arch.mask_ratio == 0.0
        m.load_state_dict(sd, strict=True)
        m.eval()
        with torch.no_grad():
            te, ve = m(data, return_embeds=True)
            assert te.shape == g["text_emb" + tag].shape
            np.testing.assert_allclose(te.numpy(), g["text_emb" + tag], atol=3e-2, rtol=3e-2)
            np.testing.assert_allclose(ve.numpy(), g["video_emb" + tag], atol=3e-2, rtol=3e-2)
            if tag == "":
                np.testing.assert_allclose(m(data, return_embeds=False).numpy(), g["sims"], atol=3e-2)
        before, emb = m.compute_video(data["video"], data["keep_ind"])          # reference return order (:94-98)
        assert before.shape == (int(g["batch"]), cfg.tokens(int(g["frames"])), cfg.embed_dim) and emb.shape == (int(g["batch"]), cfg.embed_dim)
        tb, t = m.compute_text(data["text"])
        assert tb is t and t.shape == (int(g["n_trans"]) * int(g["batch"]), cfg.embed_dim)
