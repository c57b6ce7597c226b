"""TrainStep on the GPU: (a) the flat fused AdamW against the oracle's restatement of transformers.AdamW over real kernels,
(b) the CUDA-graph replay of the whole step against kernel-by-kernel launching (same init, same batches)."""
import types

import pytest
import torch

from tvts_b200._lib import OPERAND_DTYPE

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200 import optim
from tvts_b200.synthetic import make_batch, make_state_dict
from tvts_b200.trainer import TrainStep

pytestmark = pytest.mark.gpu


def run(cfg, use_graph, n_steps=4):
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    m = m.cuda()
    opt = optim.build_reference_optimizer(m)
    for g in opt.param_groups:
        g["lr"] *= 30.0                      # visible updates within a few steps
    step = TrainStep(m, opt, cfg.temperature, torch.device("cuda"), use_graph=use_graph)
    losses = []
    try:
        for it in range(n_steps):
            data = make_batch(cfg, 2, 3, n_trans=4, seed=it)
            l1, l2 = step(data)
            losses.append((l1.item(), l2.item()))
        params = {k: p.detach().float().cpu().clone() for k, p in m.named_parameters()}
    finally:
        opt.flat.release()
    return losses, params, step


def test_graph_replay_matches_eager_steps():
    cfg = C.TINY_B_MASK
    le, pe, _ = run(cfg, False)
    lg, pg, step = run(cfg, True)
    assert step.launches_per_graph > 100
    for (a1, a2), (b1, b2) in zip(le, lg):
        assert abs(a1 - b1) < 2e-3 and abs(a2 - b2) < 2e-3, (le, lg)      # split-K / LN-gradient atomics reorder fp32 sums
    for k in pe:
        d = (pe[k] - pg[k]).abs().max().item()
        assert d < 5e-3 * max(1.0, pe[k].abs().max().item()), (k, d)
    assert le[-1][0] != le[0][0]                                          # the weights did move


def test_fused_adamw_kernel_matches_oracle_restatement():
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(5000, device="cuda")), torch.nn.Parameter(torch.randn(33, 7, device="cuda")),
          torch.nn.Parameter(torch.randn(4096, device="cuda"))]
    groups = [{"params": [ps[0], ps[2]], "lr": 1e-2, "weight_decay": 0.05}, {"params": [ps[1]], "lr": 3e-3, "weight_decay": 0.0}]
    opt = optim.AdamW(groups)
    try:
        ref = [p.detach().cpu().clone() for p in ps]
        mom = [(torch.zeros_like(r), torch.zeros_like(r)) for r in ref]
        order = [ps[0], ps[2], ps[1]]
        for t in range(1, 6):
            opt.zero_grad()
            gs = [torch.randn_like(p) for p in ps]
            for p, g in zip(ps, gs):
                if not (t == 3 and p is ps[1]):          # one parameter skips a step (grad None)
                    opt.flat.grad_view(p).copy_(g)
                    p.grad = opt.flat.grad_view(p)
            opt.step()
            for i, p in enumerate(ps):
                if p.grad is None:
                    continue
                grp = groups[0] if p is not ps[1] else groups[1]
                nsteps = t if p is not ps[1] else (t if t < 3 else t - 1)
                O.adamw_step(ref[i], gs[i].cpu(), mom[i][0], mom[i][1], nsteps, grp["lr"], weight_decay=grp["weight_decay"])
            for i, p in enumerate(ps):
                assert torch.allclose(p.detach().cpu(), ref[i], atol=2e-6, rtol=2e-6), (t, i)
                assert torch.equal(opt.flat.bf16_view(p).cpu(), p.detach().cpu().to(OPERAND_DTYPE))
    finally:
        opt.flat.release()
