"""The CPU oracle (oracle/tvts_oracle.py) must reproduce the outputs of the UNMODIFIED reference recorded in
tests/golden/*.npz by oracle/make_golden.py (the reference ships no fixtures of its own for this path)."""
import os
import sys

import numpy as np
import pytest
import torch

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200.synthetic import make_batch, make_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"tiny_B": C.TINY_B, "tiny_B_mask": C.TINY_B_MASK, "tiny_B_cap": C.TINY_B, "c1_b32": C.TVTSV2_B_32, "tiny_H": C.TINY_H}


@pytest.mark.parametrize("name", ["tiny_B", "tiny_B_mask", "tiny_B_cap", "c1_b32", "tiny_H"])
def test_oracle_matches_reference_golden(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    torch.set_num_threads(os.cpu_count())
    sd = make_state_dict(cfg, seed=1234)
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    l1, l2, (te, ve, pr), grads = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    assert abs(l1.item() - float(g["loss1"])) < 2e-5
    assert abs(l2.item() - float(g["loss2"])) < 2e-5
    np.testing.assert_allclose(te.numpy(), g["text_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(ve.numpy(), g["video_emb"], atol=2e-5, rtol=1e-4)
    if "pred_order" in g.files:
        np.testing.assert_allclose(pr.numpy(), g["pred_order"], atol=5e-5, rtol=1e-4)
    else:
        assert pr is None
    names = [str(s) for s in g["grad_names"]]
    assert set(names) == set(grads.keys())
    for k, nrm, head in zip(names, g["grad_norms"], g["grad_heads"]):
        gn = grads[k].double().norm().item()
        assert abs(gn - nrm) <= 1e-3 * nrm + 1e-7, (k, gn, nrm)
        h = grads[k].reshape(-1)[:8].double().numpy()
        np.testing.assert_allclose(h, head[: h.size], atol=1e-5 + 1e-3 * np.abs(head[: h.size]).max())


def test_v1_oracle_matches_reference_golden():
    """TVTS v1 (Conv3d tubelets, joint attention, per-tube mask, projection heads, sort head on raw tokens): oracle vs the fixture
    written by executing v1/model/model_dist_TVTS.py; the DistilBERT [CLS] vectors are an input (un-vendored text encoder)."""
    import types
    sys_path_oracle = os.path.join(os.path.dirname(GOLD), "..", "oracle")
    g = np.load(os.path.join(GOLD, "tiny_v1.npz"), allow_pickle=False)
    sys.path.insert(0, os.path.abspath(sys_path_oracle))
    from make_golden_spec import spec_state_dict
    D, heads, depth, patch, res, frames, nt, B, proj = [int(v) for v in g["dims"]]
    names = [str(s) for s in g["names"]]
    shapes = [tuple(int(x) for x in str(s).split(",")) for s in g["shapes"]]
    sd = {k: v.requires_grad_(True) for k, v in spec_state_dict(names, shapes, int(g["wseed"])).items()}
    video = torch.randn(B, frames, 3, res, res, generator=torch.Generator().manual_seed(int(g["video_seed"])))
    cfg = types.SimpleNamespace(patch=patch, width=D, heads=heads, layers=depth, sort_heads=heads, sort_depth=2, sort_ln_eps=1e-6)
    te, ve, pr = O.v1_model_forward(sd, torch.from_numpy(g["text_cls"]), video, torch.from_numpy(g["keep_ind"]), cfg)
    l1 = O.norm_softmax_loss(O.sim_matrix(ve, te), 0.05)
    l2 = O.sort_ce(pr, torch.arange(nt).repeat(B, 1))
    (l1 + l2).backward()
    assert abs(l1.item() - float(g["loss1"])) < 2e-5 and abs(l2.item() - float(g["loss2"])) < 2e-5
    np.testing.assert_allclose(te.detach().numpy(), g["text_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(ve.detach().numpy(), g["video_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(pr.detach().numpy(), g["pred_order"], atol=5e-5, rtol=1e-4)
    for k, nrm in zip([str(s) for s in g["grad_names"]], g["grad_norms"]):
        gn = sd[k].grad.double().norm().item()
        assert abs(gn - nrm) <= 1e-3 * nrm + 1e-7, (k, gn, nrm)


def test_v1_oracle_with_text_encoder_matches_reference_golden():
    """TVTS v1 end to end INCLUDING the DistilBERT text encoder (right-padded captions): oracle (with its DistilBERT restatement) vs
    the fixture written by executing v1/model/model_dist_TVTS.py around the installed transformers.DistilBertModel."""
    import v1_fixture
    g, dims, cfg, names, sd, data = v1_fixture.load()
    l1, l2, (te, ve, pr), grads = O.v1_step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg, dims.heads)
    assert abs(l1.item() - float(g["loss1"])) < 2e-5 and abs(l2.item() - float(g["loss2"])) < 2e-5
    np.testing.assert_allclose(te.numpy(), g["text_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(ve.numpy(), g["video_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(pr.numpy(), g["pred_order"], atol=5e-5, rtol=1e-4)
    assert set(grads) == set(str(s) for s in g["grad_names"])
    for k, nrm in zip([str(s) for s in g["grad_names"]], g["grad_norms"]):
        gn = grads[k].double().norm().item()
        assert abs(gn - nrm) <= 1e-3 * nrm + 1e-7, (k, gn, nrm)


def test_distilbert_restatement_matches_installed_transformers():
    """The v1 text encoder is an un-vendored dependency: the oracle's restatement is pinned against the installed implementation
    (same weights, padded and unpadded sequences)."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.DistilBertConfig(vocab_size=128, dim=128, n_layers=2, n_heads=2, hidden_dim=256, max_position_embeddings=32,
                                        dropout=0.0, attention_dropout=0.0)
    m = transformers.DistilBertModel(cfg).eval()
    sd = {"text_model." + k: v for k, v in m.state_dict().items()}
    ids = torch.randint(1, 128, (6, 12))
    mask = torch.ones_like(ids)
    mask[1, 7:] = 0
    mask[4, 3:] = 0
    ids = ids * mask
    with torch.no_grad():
        ref = m(input_ids=ids, attention_mask=mask).last_hidden_state
        out = O.distilbert_forward(sd, ids, mask, 2)
    assert (ref - out).abs().max().item() < 1e-5
