"""Model-level parity on the B200: the full TVTSv2 forward + losses + backward on the CUDA kernels against
 (a) the CPU oracle on the same seeded inputs (tiny configs, seconds on CPU) and
 (b) the committed golden fixtures written by executing the unmodified reference (tests/golden/*.npz).
Tolerances: bf16 GEMM operands / activations with fp32 accumulation -> embeddings 3e-2, losses 5e-2 at the tiny widths;
the production-width fixture (c1_b32, BASELINE.json configs[0]) is held to 2e-2 on both losses.  Token / EOT / label
indexing is exact by construction (checked bit-exactly in tests/test_kernels_gpu.py)."""
import os
import types

import numpy as np
import pytest
import torch

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200.synthetic import make_batch, make_state_dict

from tvts_b200._lib import DEFAULT_LOSS_SCALE as LOSS_SCALE

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build(cfg):
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd


def to_cuda(data):
    return {k: (v.cuda() if k != "keep_ind" else v) for k, v in data.items()}   # keep_ind stays on the host like the reference


def run_step(m, data, cfg):
    te, ve, pred = m(data)
    loss1 = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te))
    loss2 = E.sort_ce(pred, data["label"]) if pred is not None else torch.zeros((), device="cuda")
    if LOSS_SCALE == 1.0:
        (loss1 + loss2).backward()
    else:                                   # fp16-operand build (TVTS_OPERAND=fp16): backward under the static loss scale, like TrainStep
        ((loss1 + loss2) * LOSS_SCALE).backward()
        for p in m.parameters():
            if p.grad is not None:
                p.grad.div_(LOSS_SCALE)
    torch.cuda.synchronize()
    return loss1.item(), loss2.item(), te.detach().cpu(), ve.detach().cpu(), None if pred is None else pred.detach().cpu()


@pytest.mark.parametrize("cfg,batch,frames,n_trans", [(C.TINY_B, 3, 2, 4), (C.TINY_B_MASK, 2, 3, 4), (C.TINY_B, 4, 2, 1),
                                                      (C.TINY_B_MASK, 2, 8, 4)])
def test_model_matches_oracle(cfg, batch, frames, n_trans):
    m, sd = build(cfg)
    data = make_batch(cfg, batch, frames, n_trans=n_trans, seed=5)
    l1, l2, te, ve, pred = run_step(m, to_cuda(data), cfg)
    o1, o2, (ote, ove, opred), ograds = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    assert torch.allclose(te, ote, atol=3e-2, rtol=3e-2), (te - ote).abs().max()
    assert torch.allclose(ve, ove, atol=3e-2, rtol=3e-2), (ve - ove).abs().max()
    assert abs(l1 - o1.item()) < 5e-2, (l1, o1.item())
    if n_trans > 1:
        assert torch.allclose(pred, opred, atol=5e-2, rtol=5e-2), (pred - opred).abs().max()
        assert abs(l2 - o2.item()) < 5e-2, (l2, o2.item())
    else:
        assert pred is None
    got = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    assert set(got.keys()) == set(ograds.keys()), set(got.keys()) ^ set(ograds.keys())
    for k, g in ograds.items():
        rel = (got[k].double() - g.double()).norm().item() / (g.double().norm().item() + 1e-8)
        assert rel < 0.08, (k, rel)


def test_c1_against_reference_golden():
    """BASELINE.json configs[0]: TVTSv2 ViT-B/32, 2 frames, 4 clip-caption pairs; fixture from the executed reference."""
    cfg = C.TVTSV2_B_32
    g = np.load(os.path.join(GOLD, "c1_b32.npz"), allow_pickle=False)
    m, _ = build(cfg)
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    l1, l2, te, ve, pred = run_step(m, to_cuda(data), cfg)
    assert abs(l1 - float(g["loss1"])) < 2e-2, (l1, float(g["loss1"]))
    assert abs(l2 - float(g["loss2"])) < 2e-2, (l2, float(g["loss2"]))
    np.testing.assert_allclose(te.numpy(), g["text_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(ve.numpy(), g["video_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(pred.numpy(), g["pred_order"], atol=5e-2, rtol=5e-2)
    names = [str(s) for s in g["grad_names"]]
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(names) == set(got.keys())
    bad = []
    for k, nrm in zip(names, g["grad_norms"]):
        gn = got[k].double().norm().item()
        if abs(gn - nrm) > 0.05 * nrm + 1e-6:
            bad.append((k, gn, float(nrm)))
    assert not bad, bad[:8]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("wl,batch", [("c3", 2), ("c2", 2)])
def test_production_shape_matches_oracle(wl, batch):
    """The shapes bench.py times (BASELINE.json configs[2] / [1]): real ViT-B/16 (mask 0.5, N = 785) / ViT-B/32 towers, T = 8, n_trans = 4,
    a 2-pair sample: the only cases that run the 99-token space groups, the T = 8 time kernels and the deep-K split-K weight gradients.
    Losses, embeddings, order logits and EVERY parameter gradient against the CPU oracle (a few seconds of host time)."""
    w = C.WORKLOADS[wl]
    cfg = w.arch
    m, sd = build(cfg)
    data = make_batch(cfg, batch, w.frames, n_trans=w.n_trans, seed=11)
    l1, l2, te, ve, pred = run_step(m, to_cuda(data), cfg)
    o1, o2, (ote, ove, opred), ograds = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    assert torch.allclose(te, ote, atol=3e-2, rtol=3e-2), (te - ote).abs().max()
    assert torch.allclose(ve, ove, atol=3e-2, rtol=3e-2), (ve - ove).abs().max()
    assert torch.allclose(pred, opred, atol=5e-2, rtol=5e-2), (pred - opred).abs().max()
    assert abs(l1 - o1.item()) < 2e-2 and abs(l2 - o2.item()) < 2e-2, (l1, o1.item(), l2, o2.item())
    got = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    assert set(got.keys()) == set(ograds.keys()), set(got.keys()) ^ set(ograds.keys())
    worst = max(((got[k].double() - g.double()).norm().item() / (g.double().norm().item() + 1e-8), k) for k, g in ograds.items())
    print(f"{wl}: |dloss1|={abs(l1 - o1.item()):.2e} |dloss2|={abs(l2 - o2.item()):.2e} worst grad rel-L2 {worst[0]:.3f} ({worst[1]})")
    assert worst[0] < 0.08, worst


def test_missing_library_is_loud(monkeypatch):
    from tvts_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtvts_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()
