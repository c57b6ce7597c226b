"""Host logic of the flat AdamW (tvts_b200/optim.py) against the oracle's restatement of transformers.AdamW
(oracle/tvts_oracle.py:adamw_step) and the reference's parameter-group policy, with the kernel replaced by tests/emu.py."""
import types

import torch

from tvts_b200._lib import OPERAND_DTYPE

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200 import optim
from tvts_b200.synthetic import make_batch, make_state_dict
from tvts_b200.trainer import TrainStep


def test_param_group_policy_matches_reference_script():
    cfg = C.TINY_B.small(text_layers=4)
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    groups = optim.reference_param_groups(list(m.named_parameters()), text_layers=4, tune_from=3)
    names = {id(p): n for n, p in m.named_parameters()}
    g = [[names[id(p)] for p in grp["params"]] for grp in groups]
    assert "video_model.transformer.resblocks.0.timeattn.qkv.weight" in g[0] and "pred_model.head.weight" in g[0]
    assert "video_model.transformer.resblocks.0.ln_3.weight" in g[1] and "pred_model.norm.weight" in g[1]
    assert "video_model.conv1.weight" in g[2] and "video_model.temporal_embedding" in g[2] and "text_projection" in g[2]
    assert "text_model.resblocks.3.attn.in_proj_weight" in g[2] and "text_ln_final.weight" in g[3]
    assert "pred_model.type_embed" in g[0]
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert frozen and all(n.startswith("text_model.resblocks.") and int(n.split(".")[2]) < 3 for n in frozen)
    assert [grp["lr"] for grp in groups] == [1e-4, 1e-4, 1e-7, 1e-7] and [grp["weight_decay"] for grp in groups] == [0.05, 0, 0.05, 0]


def test_flat_adamw_steps_match_oracle(emu_backend):
    cfg = C.TINY_B
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)
    opt = optim.build_reference_optimizer(m)
    for g in opt.param_groups:          # large steps so that 3 updates are visible in fp32
        g["lr"] *= 100.0
    try:
        step = TrainStep(m, opt, cfg.temperature, torch.device("cpu"))
        names = {id(p): n for n, p in m.named_parameters()}
        ref = {n: p.detach().clone() for n, p in m.named_parameters()}
        mom = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in ref.items()}
        nsteps = {n: 0 for n in ref}
        for it, n_trans in enumerate([4, 1, 4]):        # the caption batch leaves pred_model without gradients
            data = make_batch(cfg, 2, 2, n_trans=n_trans, seed=it)
            E.WEIGHTS.clear()
            step(data)
            got = {names[id(p)]: (None if p.grad is None else p.grad.clone() / step.loss_scale) for p in opt.flat.params}   # (fp16 build: scaled)
            if n_trans == 1:
                assert all(g is None for k, g in got.items() if k.startswith("pred_model"))
            for gi, grp in enumerate(opt.param_groups):
                for p in grp["params"]:
                    n = names[id(p)]
                    if got[n] is None:
                        continue
                    nsteps[n] += 1
                    O.adamw_step(ref[n], got[n], mom[n][0], mom[n][1], nsteps[n], grp["lr"], weight_decay=grp["weight_decay"])
            for p in opt.flat.params:
                n = names[id(p)]
                assert torch.allclose(p.detach(), ref[n], atol=1e-6, rtol=1e-5), (it, n, (p.detach() - ref[n]).abs().max())
                bf = opt.flat.bf16_view(p)
                assert torch.equal(bf, p.detach().to(OPERAND_DTYPE)), n
        assert any(v == 2 for k, v in nsteps.items() if k.startswith("pred_model")) and nsteps["video_model.proj"] == 3
    finally:
        opt.flat.release()


def test_flatstate_active_mask_roundtrip():
    ps = [torch.nn.Parameter(torch.randn(10)), torch.nn.Parameter(torch.randn(3, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = optim.AdamW([{"params": ps, "lr": 1e-3}])
    try:
        fs = opt.flat
        ps[0].grad = fs.grad_view(ps[0])
        ps[2].grad = fs.grad_view(ps[2])
        mask = fs.active_mask()
        assert mask == [True, False, True]
        fs.zero_grad()
        assert fs.active_mask() == [False, False, False]
        fs.set_active(mask)
        assert fs.active_mask() == mask and ps[0].grad.data_ptr() == fs.grad_view(ps[0]).data_ptr()
        fs.set_active([False, True, False])
        assert ps[0].grad is None and ps[1].grad is not None and ps[2].grad is None
    finally:
        opt.flat.release()
