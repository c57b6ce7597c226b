"""TVTS v1 (BASELINE.json configs[4]) host logic on the torch emulation of the kernels: tubelet video tower with per-tube masks,
the DistilBERT text tower (post-LN, key-padding mask), ReLU/Linear projection heads, sort head on the raw tokens -- against the
fixture written by the executed reference (tests/golden/tiny_v1_full.npz) and the oracle's gradients; plus the v1 trainer loop."""
import os
import subprocess
import sys
import types

import numpy as np
import torch

import tvts_oracle as O
import v1_fixture
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200 import modules_v1 as V1
from tvts_b200 import optim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(dims):
    text_model = V1.DistilBertShell(vocab_size=dims.vocab, dim=dims.D, n_layers=2, n_heads=dims.heads, hidden_dim=2 * dims.D,
                                    max_position_embeddings=32)
    video_model = V1.VisionTransformer(img_size=dims.res, patch_size=dims.patch, embed_dim=dims.D, depth=dims.depth, num_heads=dims.heads,
                                       num_frames=dims.frames)
    return V1.TVTS(types.SimpleNamespace(local_rank=0), {"num_frames": dims.frames}, {"model": "distilbert-base-uncased", "pretrained": True},
                   projection_dim=dims.proj, text_model=text_model, video_model=video_model, sort_heads=dims.heads)


def test_v1_model_matches_reference_golden_and_oracle(emu_backend):
    E.WEIGHTS.clear()
    g, dims, cfg, names, sd, data = v1_fixture.load()
    m = build(dims)
    assert [k for k, _ in m.named_parameters()] == names          # names AND enumeration order of the reference (HF DistilBERT included)
    m.load_state_dict(sd, strict=True)
    te, ve, pred = m(data)
    l1 = M.NormSoftmaxLoss(0.05)(M.sim_matrix(ve, te))
    l2 = E.sort_ce(pred, data["label"], 2.0)
    (l1 + l2).backward()
    assert abs(l1.item() - float(g["loss1"])) < 2e-2 and abs(l2.item() - float(g["loss2"])) < 2e-2
    np.testing.assert_allclose(te.detach().numpy(), g["text_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(ve.detach().numpy(), g["video_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(pred.detach().numpy(), g["pred_order"], atol=5e-2, rtol=5e-2)
    _, _, _, ograds = O.v1_step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg, dims.heads)
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ograds)
    for k, gr in ograds.items():
        if gr.norm().item() < 1e-6:          # k_lin.bias: exactly-zero true gradient (softmax shift invariance)
            assert got[k].norm().item() < 1e-3
            continue
        rel = (got[k].double() - gr.double()).norm().item() / gr.double().norm().item()
        assert rel < 0.08, (k, rel)
    # reference return conventions (model_dist_TVTS.py:119-141)
    tb, t = m.compute_text(data["text"])
    assert tb.shape == (dims.nt * dims.B, dims.D) and t.shape == (dims.nt * dims.B, dims.proj)
    tokens, v = m.compute_video(data["video"], data["keep_ind"])
    assert tokens.shape == (dims.B, 1 + (dims.frames // 2) * data["keep_ind"].shape[-1], dims.D) and v.shape == (dims.B, dims.proj)


class FakeHFTokenizer:
    """return_tensors='pt', padding=True|'max_length', truncation, max_length: right-padded ids + attention_mask, like HF tokenizers."""

    def __init__(self, vocab):
        self.vocab, self.calls = vocab, []

    def __call__(self, texts, return_tensors="pt", padding=True, truncation=True, max_length=50):
        self.calls.append((list(texts), padding, max_length))
        lens = [3 + (len(t) * 7) % 9 for t in texts]
        Lc = max_length if padding == "max_length" else max(lens)
        g = torch.Generator().manual_seed(len(self.calls))
        ids = torch.randint(1, self.vocab, (len(texts), Lc), generator=g)
        mask = (torch.arange(Lc)[None, :] < torch.tensor(lens)[:, None]).long()
        return {"input_ids": ids * mask, "attention_mask": mask}


class Loader:
    dataset_name, batch_size = "YTTemporal", 2

    def __init__(self, batches):
        self.batches = batches
        self.train_sampler = types.SimpleNamespace(set_epoch=lambda e: None)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def test_v1_trainer_loop_tokenises_like_the_reference_and_steps(emu_backend):
    from tvts_b200.trainer import Trainer_TVTS
    E.WEIGHTS.clear()
    g, dims, cfg, names, sd, data = v1_fixture.load()
    m = build(dims)
    m.load_state_dict(sd, strict=True)
    opt = optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0)   # dist-yt-pt.json:44-54
    tok = FakeHFTokenizer(dims.vocab)
    captions = [[f"clip{t} sample{b} {'x' * (t + 2 * b)}" for b in range(dims.B)] for t in range(dims.nt)]
    batches = [dict(video=data["video"] + i, keep_ind=data["keep_ind"], label=data["label"], text=captions) for i in range(2)]
    args = types.SimpleNamespace(rank=0, local_rank=0, world_size=1, schedule=[2])
    try:
        tr = Trainer_TVTS(args, m, M.NormSoftmaxLoss(0.05), [], opt, {"trainer": {"epochs": 2}}, [Loader(batches)], tokenizer=tok,
                          use_graph=False)
        w0 = m.text_model.transformer.layer[0].ffn.lin1.weight.detach().clone()
        hist = tr.train()
        assert len(hist) == 2 and hist[0]["loss_0"] > 0
        assert len(tok.calls) == 4 and tok.calls[0][0][:3] == [captions[0][0], captions[0][1], captions[1][0]]      # clip-major
        assert tok.calls[0][1] is True and tok.calls[0][2] == 50                                                     # trainer.py:130-131
        assert not torch.equal(w0, m.text_model.transformer.layer[0].ffn.lin1.weight.detach())                        # DistilBERT is trained
        assert opt.param_groups[0]["lr"] == 1e-3 * 0.1                        # absolute schedule: base_lr x 0.1 once epoch >= 2
        assert all(s == 4 for s in opt.sync_steps())
    finally:
        opt.flat.release()


def test_v1_drop_in_paths_resolve():
    code = ("import model.model_dist_TVTS as a, model.video_encoder as v, model.sort_transformer as s, model.loss as l, trainer.trainer as t;"
            "from trainer import Trainer_TVTS; import model.metric as mm; assert mm.t2v_metrics and mm.v2t_metrics;"
            "assert a.TVTS and a.sim_matrix and v.VisionTransformer and s.SortTransformer and l.NormSoftmaxLoss and t.AllGather_multi")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tvts_b200", "dropin_v1") + os.pathsep + ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")


def test_v1_epoch_length_is_the_longest_loader(emu_backend):
    """v1/trainer/trainer.py:50-54: `len_epoch = max(len(x) for x in data_loader)` (v2 takes the YT loader instead): the longest loader
    drives an epoch, the others are cycled (round-1 ADVICE)."""
    from tvts_b200.trainer import Trainer_TVTS
    E.WEIGHTS.clear()
    g, dims, cfg, names, sd, data = v1_fixture.load()
    m = build(dims)
    m.load_state_dict(sd, strict=True)
    opt = optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.0)
    batch = dict(video=data["video"], keep_ind=data["keep_ind"], label=data["label"], text=data["text"])
    try:
        tr = Trainer_TVTS(types.SimpleNamespace(rank=0, local_rank=0, world_size=1, schedule=[]), m, M.NormSoftmaxLoss(0.05), [], opt,
                          {"trainer": {"epochs": 1}}, [Loader([batch]), Loader([batch, batch, batch])], use_graph=False)
        assert tr.len_epoch == 3
        hist = tr.train()
        assert max(opt.sync_steps()) == 6            # 3 iterations x 2 loaders, the short loader cycled
    finally:
        opt.flat.release()
