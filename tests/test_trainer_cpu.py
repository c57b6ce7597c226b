"""Epoch loop of Trainer_TVTSv2_* (tvts_b200/trainer.py) on the torch emulation of the kernels: loader interleaving, clip-major
caption flattening + tokenisation, one optimizer step per loader batch, milestone decay, validation metrics."""
import types

import torch

from tvts_b200 import config as C
from tvts_b200 import metrics as MT
from tvts_b200 import modules as M
from tvts_b200 import optim
from tvts_b200.synthetic import make_batch, make_state_dict, make_tokens
from tvts_b200.trainer import Trainer_TVTSv2_B_16


class FakeLoader:
    def __init__(self, name, batches, batch_size):
        self.dataset_name, self.batches, self.batch_size = name, batches, batch_size
        self.epochs_seen = []
        self.train_sampler = types.SimpleNamespace(set_epoch=self.epochs_seen.append)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def captions(batch, n_trans, tag):
    return [[f"{tag} clip{t} sample{b}" for b in range(batch)] for t in range(n_trans)]       # [clip][sample], like the datasets


def test_epoch_loop_interleaves_loaders_and_decays_lr(emu_backend):
    cfg = C.TINY_B
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    opt = optim.build_reference_optimizer(m)
    seen = []

    def tokenizer(texts, truncate=True):
        seen.append(list(texts))
        return make_tokens(cfg, len(texts), seed=len(seen))

    def yt_batch(i):
        b = make_batch(cfg, 2, 2, n_trans=4, seed=i)
        b["text"] = captions(2, 4, f"yt{i}")
        return b

    def web_batch(i):
        b = make_batch(cfg, 3, 2, n_trans=1, seed=50 + i)
        b["text"] = captions(3, 1, f"web{i}")
        del b["label"]
        return b

    yt = FakeLoader("YTTemporal", [yt_batch(i) for i in range(3)], 2)
    web = FakeLoader("WebVid", [web_batch(i) for i in range(2)], 3)          # shorter: must be cycled
    val = [FakeLoader("MSRVTT", [yt_batch(9)], 2)]
    args = types.SimpleNamespace(rank=0, local_rank=0, world_size=1, schedule=[1])
    cfgd = {"trainer": {"epochs": 2}}
    try:
        tr = Trainer_TVTSv2_B_16(args, m, M.NormSoftmaxLoss(0.05), [MT.t2v_metrics, MT.v2t_metrics], opt, cfgd, [yt, web],
                                 valid_data_loader=val, tokenizer=tokenizer, use_graph=False)
        lr0 = [g["lr"] for g in opt.param_groups]
        w0 = m.video_model.proj.detach().clone()
        hist = tr.train()
        assert len(hist) == 2 and tr.len_epoch == 3
        # 3 YT + 3 WebVid steps per epoch (WebVid cycled), + 1 validation batch per epoch
        train_calls = [t for t in seen if not t[0].startswith("yt9")]
        assert len(train_calls) == 12
        assert train_calls[0][:3] == ["yt0 clip0 sample0", "yt0 clip0 sample1", "yt0 clip1 sample0"]      # clip-major flattening
        assert train_calls[5][0].startswith("web0")                                                        # third WebVid step wrapped around
        assert yt.epochs_seen == [1, 2] and web.epochs_seen == [1, 2]
        assert [g["lr"] for g in opt.param_groups] == [l * 0.1 for l in lr0]                                 # milestone at epoch 1 only
        assert all(k in hist[0] for k in ("loss_0", "loss_1", "val_0_t2v_metrics_R1", "val_loss_0"))
        assert hist[0]["loss_0"] > 0 and hist[0]["loss_1"] > 0
        assert not torch.equal(w0, m.video_model.proj.detach())
        steps = {id(p): s for s, p in zip(opt.steps, opt.flat.params)}
        assert steps[id(m.video_model.proj)] == 12                         # every loader batch of both epochs
        assert steps[id(m.pred_model.head.weight)] == 6                    # transcript (YT) batches only: WebVid steps skip the sort head
    finally:
        opt.flat.release()
