"""Host-side orchestration (tvts_b200/engine.py + modules.py) against the oracle, with the native ops replaced by the
torch restatement of tests/emu.py -- this checks wiring, layouts, the hand-written backward and parameter naming
without a GPU.  The same comparison with the REAL kernels runs under -m gpu (tests/test_model_gpu.py)."""
import types

import pytest
import torch

from tvts_b200._lib import OPERAND_DTYPE

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200.synthetic import make_batch, make_state_dict


def build(cfg):
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)          # parameter names / shapes are the reference's
    return m, sd


def run_step(m, data, cfg):
    E.WEIGHTS.clear()
    te, ve, pred = m(data)
    loss1 = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te))
    loss2 = E.sort_ce(pred, data["label"]) if pred is not None else torch.zeros(())
    (loss1 + loss2).backward()
    return loss1.detach(), loss2.detach(), te.detach(), ve.detach(), None if pred is None else pred.detach()


@pytest.mark.parametrize("cfg,batch,frames,n_trans", [(C.TINY_B, 3, 2, 4), (C.TINY_B_MASK, 2, 3, 4), (C.TINY_B, 4, 2, 1),
                                                      (C.TINY_B_MASK, 1, 12, 2),      # the shipped 12-frame clips, 2 transcripts
                                                      (C.TINY_B, 2, 1, 3)])           # single frame (4-D video path: T = 1)
def test_engine_matches_oracle(emu_backend, cfg, batch, frames, n_trans):
    torch.manual_seed(0)
    m, sd = build(cfg)
    data = make_batch(cfg, batch, frames, n_trans=n_trans, seed=5)
    l1, l2, te, ve, pred = run_step(m, data, cfg)
    o1, o2, (ote, ove, opred), ograds = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    # bf16 operands / activations: embeddings to ~1e-2 relative, losses to a few 1e-2 absolute at these tiny widths
    assert torch.allclose(te, ote, atol=3e-2, rtol=3e-2), (te - ote).abs().max()
    assert torch.allclose(ve, ove, atol=3e-2, rtol=3e-2), (ve - ove).abs().max()
    assert abs(l1.item() - o1.item()) < 5e-2
    if n_trans > 1:
        assert torch.allclose(pred, opred, atol=5e-2, rtol=5e-2)
        assert abs(l2.item() - o2.item()) < 5e-2
    else:
        assert pred is None
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got.keys()) == set(ograds.keys()), set(got.keys()) ^ set(ograds.keys())
    worst = 0.0
    for k, g in ograds.items():
        num = (got[k].double() - g.double()).norm().item()
        den = g.double().norm().item() + 1e-8
        worst = max(worst, num / den)
        assert num / den < 0.08, (k, num / den)
    print("worst relative grad error", worst)


def test_unused_sort_head_gets_no_grad(emu_backend):
    """caption batches (n_trans == 1) leave pred_model out of the graph (find_unused_parameters semantics)."""
    cfg = C.TINY_B
    m, _ = build(cfg)
    data = make_batch(cfg, 2, 2, n_trans=1, seed=1)
    run_step(m, data, cfg)
    assert all(p.grad is None for k, p in m.named_parameters() if k.startswith("pred_model"))
    assert all(p.grad is not None for k, p in m.named_parameters() if not k.startswith("pred_model"))


def test_weight_cache_rejects_recycled_parameter_ids(emu_backend):
    """id() values are recycled after a model is garbage-collected: a cache entry keyed by a dead parameter's id must not be
    served for a new parameter that happens to get the same id / storage / version (this produced NaNs between GPU tests)."""
    import weakref
    cache = E._WeightCache()
    old = torch.nn.Parameter(torch.randn(4, 4))
    stale = cache.get(old)
    new = torch.nn.Parameter(torch.randn(8, 8))
    cache._c[id(new)] = (new._version, new.data_ptr(), stale, weakref.ref(old))     # what id reuse leaves behind
    got = cache.get(new)
    assert got.shape == new.shape and torch.equal(got, new.detach().to(OPERAND_DTYPE))
    assert cache.get(new) is got                                                       # now a genuine hit
    with torch.no_grad():
        new.add_(1.0)                                                                   # version bump -> re-cast
    assert torch.equal(cache.get(new), new.detach().to(OPERAND_DTYPE))


def test_validate_runs_forward_only(emu_backend):
    from tvts_b200.trainer import validate
    cfg = C.TINY_B
    m, sd = build(cfg)
    batches = [make_batch(cfg, 6, 2, n_trans=4, seed=9)]
    res = validate(m, batches, device=torch.device("cpu"))
    assert set(res) == {"t2v_metrics", "v2t_metrics", "order_acc"}
    assert 0.0 <= res["t2v_metrics"]["R1"] <= 100.0 and res["order_acc"] is not None
    assert all(p.grad is None for p in m.parameters())


@pytest.mark.parametrize("cfg,name", [(C.TINY_H, "tiny_H"), (C.TINY_H640, "tiny_H640")])
def test_h14_engine_matches_reference_golden_and_oracle(emu_backend, cfg, name):
    """TVTSv2_H_14 semantics (model_dist_TVTSv2_ViT_H_14.py, video_encoder_ViT_H_14.py): 14x14 patches through the padded patch-embed
    GEMM, head dim 80, exact GELU, ln_post on the CLS row only, sort head over the patch tokens -- against the fixture written by the
    executed reference and the oracle's gradients; parameter enumeration order = the reference's."""
    import os
    import numpy as np
    E.WEIGHTS.clear()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    m = M.TVTSv2_H_14(types.SimpleNamespace(local_rank=0), arch=cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    l1, l2, te, ve, pred = run_step(m, data, cfg)
    assert abs(l1.item() - float(g["loss1"])) < 2e-2 and abs(l2.item() - float(g["loss2"])) < 2e-2
    np.testing.assert_allclose(te.numpy(), g["text_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(ve.numpy(), g["video_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(pred.numpy(), g["pred_order"], atol=5e-2, rtol=5e-2)
    tokens, pooled = m.compute_video(data["video"], data["keep_ind"])
    assert tokens.shape == (int(g["batch"]), cfg.tokens(int(g["frames"])) - 1, cfg.embed_dim) and pooled.shape == ve.shape
    _, _, _, ograds = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ograds)
    for k, gr in ograds.items():
        rel = (got[k].double() - gr.double()).norm().item() / (gr.double().norm().item() + 1e-8)
        assert rel < 0.08, (k, rel)
    ref_order = [str(s) for s in g["grad_names"]]
    assert [k for k, _ in m.named_parameters() if k in set(ref_order)] == ref_order


def test_uint8_clips_take_the_fused_input_stage(emu_backend):
    """uint8 crops fed to the model give EXACTLY what the reference's CPU transform (x/255, then (x-mean)/std in fp32:
    video_transforms/video_transform.py:24-76,627-650) followed by the float path gives -- forward and gradients."""
    cfg = C.TINY_B_MASK
    data = make_batch(cfg, 2, 3, n_trans=4, seed=7)
    u8 = torch.randint(0, 256, data["video"].shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    mean = torch.tensor(cfg.input_mean)[None, None, :, None, None]
    std = torch.tensor(cfg.input_std)[None, None, :, None, None]
    ref_video = u8.float().div(255).sub(mean).div(std)
    outs = []
    for video in (u8, ref_video):
        torch.manual_seed(0)
        m, _ = build(cfg)
        E.WEIGHTS.clear()
        l1, l2, te, ve, pred = run_step(m, dict(data, video=video), cfg)
        outs.append((l1, l2, ve, pred, m.video_model.conv1.weight.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
