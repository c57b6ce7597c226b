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
            "assert a.TVTSv2_B_16 and b.TVTSv2_B_32 and a.sim_matrix and l.NormSoftmaxLoss and t.AllGather_multi and s.SortTransformer and v.VisionTransformer")
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
