"""Multi-rank semantics of the step on 2 gloo ranks (CPU, kernels replaced by tests/emu.py): the fused embedding gather has
AllGather_multi's forward / local-slice backward (v2/trainer/trainer.py:41-57), the global loss equals the single-process loss on
the concatenated batch, and averaged parameter gradients are exactly 1/W of the single-process ones."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        import tvts_oracle as O
        from tvts_b200 import config as C
        from tvts_b200 import engine as E
        from tvts_b200 import modules as M
        from tvts_b200 import optim
        from tvts_b200.synthetic import make_batch, make_state_dict
        from tvts_b200.trainer import AllGather_multi, TrainStep, gather_embeddings
        emu.install()
        torch.set_num_threads(2)
        # --- AllGather_multi: forward concatenates in rank order, backward returns the local slice
        x = (torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * rank).requires_grad_(True)
        y = AllGather_multi.apply(x, world, types.SimpleNamespace(rank=rank))
        assert y.shape == (3 * world, 2) and torch.equal(y[3 * rank: 3 * rank + 3], x.detach())
        w = torch.arange(3 * world * 2, dtype=torch.float32).view(3 * world, 2)
        (y * w).sum().backward()
        assert torch.equal(x.grad, w[3 * rank: 3 * rank + 3])
        # --- one step on a 2-rank global batch vs the oracle on the concatenated batch
        cfg = C.TINY_B
        Bl, T, nt = 2, 2, 4
        sd = make_state_dict(cfg, seed=1234)
        m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0, rank=rank), arch=cfg)
        m.load_state_dict(sd, strict=True)
        opt = optim.build_reference_optimizer(m)
        E.WEIGHTS.clear()
        step = TrainStep(m, None, cfg.temperature, torch.device("cpu"))
        step.optimizer = types.SimpleNamespace(flat=opt.flat, zero_grad=opt.zero_grad, step=lambda *a: None, launch=lambda *a: None)   # keep the weights fixed
        local = make_batch(cfg, Bl, T, n_trans=nt, seed=3, rank=rank)
        step.overlap = True                                # the opt-in bucketed all-reduce first, then the single all-reduce
        l1, l2 = step(local)
        grads = {k: p.grad.clone() / step.loss_scale for k, p in m.named_parameters() if p.grad is not None}   # (fp16 build: loss-scaled)
        # bucketed overlap (the default when W > 1; the first step above ran it, this one runs the single all-reduce): buckets are sent from
        # engine.GRAD_READY hooks while the backward runs -- same sums, so the averaged gradients must be IDENTICAL, every bucket must
        # have been sent by its hook, and the buckets must cover the arena exactly once
        assert step.overlap and E.GRAD_READY is None
        cover = sorted(r for runs in step._ranges.values() for r in runs)
        assert cover[0][0] == 0 and cover[-1][1] == opt.flat.total and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        assert {("text",), ("sort",), ("rest",), ("video_block", 0)} <= set(step._ranges)
        assert step._sent == set(step._ranges) - {("rest",)}
        step.overlap = False
        step(local)
        for k, p in m.named_parameters():
            if p.grad is not None:
                assert torch.equal(p.grad / step.loss_scale, grads[k]), k
        if rank == 0:
            parts = [make_batch(cfg, Bl, T, n_trans=nt, seed=3, rank=r) for r in range(world)]
            video = torch.cat([p["video"] for p in parts])
            keep = torch.cat([p["keep_ind"] for p in parts])
            label = torch.cat([p["label"] for p in parts])
            # clip-major text rows: index t*B + b over the GLOBAL batch
            text = torch.cat([torch.stack([p["text"][t * Bl:(t + 1) * Bl] for p in parts]).reshape(-1, cfg.context) for t in range(nt)])
            trainable = {k for k, p in m.named_parameters() if p.requires_grad}
            o1, o2, _, og = O.step_with_grads(sd, text, video, keep, label, cfg, trainable=trainable)
            out["loss1"] = (l1.item(), o1.item())
            worst = 0.0
            for k, g in og.items():
                if k.startswith("pred_model") or k not in grads:
                    continue
                # InfoNCE part: local-slice backward + averaging => 1/W of the global gradient; the video/text towers also carry the
                # LOCAL sort-loss gradient averaged over ranks = the global mean-CE gradient, so the whole thing is 1/W * global for
                # the contrastive part and exactly global for the sort part; compare on parameters only the contrastive loss reaches
                if not k.startswith("text_"):
                    continue
                ref = g / world
                rel = (grads[k] - ref).norm().item() / (ref.norm().item() + 1e-12)
                worst = max(worst, rel)
            out["worst_text_grad_rel"] = worst
        opt.flat.release()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_single_process_oracle():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = out["loss1"]
    assert abs(a - b) < 5e-2, (a, b)
    assert out["worst_text_grad_rel"] < 0.08, out["worst_text_grad_rel"]


def _worker_v1(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        import tvts_oracle as O
        import v1_fixture
        from test_v1_cpu import build
        from tvts_b200 import engine as E
        from tvts_b200 import optim
        from tvts_b200.trainer import TrainStep
        emu.install()
        torch.set_num_threads(2)
        g, dims, cfg, names, sd, data = v1_fixture.load()          # global batch of 2 pairs -> 1 pair per rank
        assert dims.B == world
        m = build(dims)
        m.load_state_dict(sd, strict=True)
        opt = optim.AdamW([p for p in m.parameters()], lr=1e-4, weight_decay=0.0)
        E.WEIGHTS.clear()
        step = TrainStep(m, None, 0.05, torch.device("cpu"))
        step.optimizer = types.SimpleNamespace(flat=opt.flat, zero_grad=opt.zero_grad, step=lambda *a: None, launch=lambda *a: None)   # keep the weights fixed
        rows = torch.tensor([t * dims.B + rank for t in range(dims.nt)])           # this rank's captions, clip-major
        local = {"video": data["video"][rank:rank + 1], "keep_ind": data["keep_ind"][rank:rank + 1], "label": data["label"][rank:rank + 1],
                 "text": {k: v[rows] for k, v in data["text"].items()}}
        l1, l2 = step(local)
        grads = {k: p.grad.clone() / step.loss_scale for k, p in m.named_parameters() if p.grad is not None}   # (fp16 build: loss-scaled)
        if rank == 0:
            o1, o2, _, og = O.v1_step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg, dims.heads)
            out["loss1"] = (l1.item(), o1.item())
            worst = 0.0
            for k, gr in og.items():                      # parameters only the contrastive loss reaches: 1/W of the global gradient
                if not (k.startswith("text_model.") or k.startswith("txt_proj.")) or gr.norm().item() < 1e-6:
                    continue
                ref = gr / world
                worst = max(worst, (grads[k] - ref).norm().item() / (ref.norm().item() + 1e-12))
            out["worst_text_grad_rel"] = worst
        opt.flat.release()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_v1_step_matches_single_process_oracle():
    """TVTS v1 (DistilBERT captions are clip-major across the GLOBAL batch) on 2 gloo ranks vs the oracle on the whole batch."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_v1, args=(2, port, out), nprocs=2, join=True)
    a, b = out["loss1"]
    assert abs(a - b) < 5e-2, (a, b)
    assert out["worst_text_grad_rel"] < 0.08, out["worst_text_grad_rel"]


def _worker_pipelined(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        from tvts_b200 import config as C
        from tvts_b200 import engine as E
        from tvts_b200 import modules as M
        from tvts_b200 import optim
        from tvts_b200.synthetic import make_batch, make_state_dict
        from tvts_b200.trainer import TrainStep
        emu.install()
        torch.set_num_threads(2)
        cfg = C.TINY_B
        sd = make_state_dict(cfg, seed=1234)
        batches = [make_batch(cfg, 2, 2, n_trans=4, seed=10 + i, rank=rank) for i in range(3)]
        results = []
        for pipelined in (False, True):
            m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0, rank=rank), arch=cfg)
            m.load_state_dict(sd, strict=True)
            opt = optim.build_reference_optimizer(m)
            for g in opt.param_groups:
                g["lr"] = g["lr"] * 100.0                   # make three steps move the weights visibly
            E.WEIGHTS.clear()
            step = TrainStep(m, opt, cfg.temperature, torch.device("cpu"))
            step.pipelined = pipelined
            step.PIPELINE_BUCKETS = 3
            losses = [tuple(x.item() for x in step(b)) for b in batches]
            results.append((losses, {k: v.detach().clone() for k, v in m.state_dict().items()}, opt.sync_steps()[:], float(opt.scale_state[0])))
            if pipelined and opt.dynamic_scale:
                # a non-finite LOCAL gradient on one rank only: the flag travels, BOTH ranks skip the step and halve the scale
                w0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
                s0 = float(opt.scale_state[0])
                orig = opt.launch_check

                def poisoned():
                    if rank == 1:
                        opt.flat.g[5] = float("inf")
                    orig()
                opt.launch_check = poisoned
                step(batches[0])
                opt.launch_check = orig
                assert all(torch.equal(v, w0[k]) for k, v in m.state_dict().items()), "a rank updated its weights in a step the other rank skipped"
                assert float(opt.scale_state[0]) == s0 * 0.5 and opt.skipped_steps == 1
                out[f"skip{rank}"] = True
            opt.flat.release()
        (l_a, w_a, st_a, sc_a), (l_b, w_b, st_b, sc_b) = results
        assert l_a == l_b, (l_a, l_b)
        assert st_a == st_b and sc_a == sc_b
        for k in w_a:
            assert torch.equal(w_a[k], w_b[k]), k
        out[f"ok{rank}"] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_pipelined_allreduce_adamw_matches_single_allreduce():
    """TVTS_PIPELINED_ADAMW: bucketed all-reduce with the AdamW update of bucket i under the all-reduce of bucket i + 1 -- identical
    losses, weights, step counters and loss scale to one all-reduce followed by one AdamW launch, over three steps on 2 ranks; and a
    non-finite gradient on ONE rank makes BOTH ranks skip the step (dynamic loss scale: the finite flag is MAX-reduced)."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + ((os.getpid() + 977) % 2000)
    mp.spawn(_worker_pipelined, args=(2, port, out), nprocs=2, join=True)
    assert out.get("ok0") and out.get("ok1"), dict(out)

