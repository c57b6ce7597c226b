"""The CUDA-graph path of TrainStep (capture per input signature, static input buffers, replay, shared memory pool, gradient-attachment
restore, bf16-shadow refresh, prefetch / staging) driven on the CPU with stand-ins for torch.cuda's stream / event / graph objects:
`capture` runs the body once for its host-side bookkeeping and then rolls the numeric state back (a real capture records kernels
without executing them), `replay` re-executes the captured body on the static buffers.  The kernels are the torch restatements
(tests/emu.py).  This is a control-flow test of code that otherwise only runs on a GPU; the real thing is tests/test_trainstep_gpu.py."""
import contextlib
import types

import pytest
import torch

from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200 import optim
from tvts_b200 import trainer as TR
from tvts_b200.synthetic import make_batch, make_state_dict


class FakeStream:
    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass


class FakeEvent:
    def __init__(self, **k):
        pass

    def record(self, s=None):
        pass

    def synchronize(self):
        pass


class FakeGraph:
    captured_pools = []

    def __init__(self):
        self.fn, self.outs = None, None

    def pool(self):
        return ("pool-of", id(self))

    def reset(self):
        self.fn = None

    def replay(self):
        new = self.fn()
        for dst, src in zip(self.outs, new):
            dst.copy_(src)


@pytest.fixture
def fake_cuda(monkeypatch, emu_backend):
    state = types.SimpleNamespace(capturing=None, step=None)

    @contextlib.contextmanager
    def graph_ctx(g, pool=None):
        FakeGraph.captured_pools.append(pool)
        fs = state.step.optimizer.flat
        opt = state.step.optimizer
        snap = (fs.p.clone(), fs.m.clone(), fs.v.clone(), fs.bf.clone(), list(opt.steps), opt.steps_dev.clone(), opt.scale_state.clone())
        state.capturing = g
        try:
            yield
        finally:
            state.capturing = None
            fs.p.copy_(snap[0]); fs.m.copy_(snap[1]); fs.v.copy_(snap[2]); fs.bf.copy_(snap[3])      # a capture executes nothing
            opt.steps[:] = snap[4]
            opt.steps_dev.copy_(snap[5]); opt.scale_state.copy_(snap[6])       # (dynamic loss scaling keeps its counters on the device)

    orig_body = TR.TrainStep._body

    def body(self, data, optimizer_launch_only=False, skip_optimizer=False):
        out = orig_body(self, data, optimizer_launch_only=optimizer_launch_only, skip_optimizer=skip_optimizer)
        if state.capturing is not None:
            g = state.capturing
            g.fn = lambda: orig_body(self, data, optimizer_launch_only=True)
            g.outs = out
        return out

    monkeypatch.setattr(TR.TrainStep, "_body", body)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", graph_ctx)
    FakeGraph.captured_pools = []
    return state


def build(cfg, lr_scale=100.0):
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    opt = optim.build_reference_optimizer(m)
    for g in opt.param_groups:
        g["lr"] *= lr_scale
    return m, opt


def test_graph_path_matches_eager_steps_across_two_signatures(fake_cuda):
    """transcript batches (both losses) alternate with caption batches (no sort-head gradients) -> two graphs sharing one memory pool;
    the replayed steps must leave the same weights and losses as the eager steps."""
    cfg = C.TINY_B
    batches = [make_batch(cfg, 2, 2, n_trans=(4 if i % 2 == 0 else 1), seed=30 + i) for i in range(5)]
    results = []
    for use_graph in (False, True):
        E.WEIGHTS.clear()
        m, opt = build(cfg)
        try:
            step = TR.TrainStep(m, opt, cfg.temperature, torch.device("cpu"), use_graph=use_graph)
            fake_cuda.step = step
            losses = []
            for b in batches:
                l1, l2 = step(b)
                losses.append((l1.item(), l2.item()))
            results.append((losses, {k: p.detach().clone() for k, p in m.named_parameters()}, list(opt.steps), step))
        finally:
            opt.flat.release()
    (le, pe, se, _), (lg, pg, sg, gstep) = results
    assert len(gstep._graphs) == 2
    assert FakeGraph.captured_pools[0] is None and FakeGraph.captured_pools[1] is not None       # the 2nd signature shares the 1st pool
    assert se == sg                                                                               # per-tensor step counters (sort head: 3 of 5)
    for a, b in zip(le, lg):
        assert abs(a[0] - b[0]) < 1e-5 and abs(a[1] - b[1]) < 1e-5
    for k in pe:
        assert torch.allclose(pe[k], pg[k], atol=1e-6, rtol=1e-5), k


def test_graph_replay_sees_weights_edited_outside_the_optimizer(fake_cuda):
    """load_state_dict / checkpoint resume between replays: the operand copies the graph reads are re-cast before the next replay."""
    cfg = C.TINY_B
    E.WEIGHTS.clear()
    m, opt = build(cfg, lr_scale=1.0)
    try:
        step = TR.TrainStep(m, opt, cfg.temperature, torch.device("cpu"), use_graph=True)
        fake_cuda.step = step
        data = make_batch(cfg, 2, 2, n_trans=4, seed=3)
        step(data)
        other = make_state_dict(cfg, seed=77)
        m.load_state_dict(other, strict=True)                       # in-place copies into the arena views
        l1, l2 = step(data)
        E.WEIGHTS.clear()
        m2, opt2 = build(cfg, lr_scale=1.0)
        opt.flat.release()
        m2.load_state_dict(other, strict=True)
        ref = TR.TrainStep(m2, opt2, cfg.temperature, torch.device("cpu"))(data)
        assert abs(l1.item() - ref[0].item()) < 1e-5 and abs(l2.item() - ref[1].item()) < 1e-5
    finally:
        opt.flat.release()
        if "opt2" in locals():
            opt2.flat.release()


def test_prefetch_staging_feeds_the_graph(fake_cuda):
    cfg = C.TINY_B
    E.WEIGHTS.clear()
    m, opt = build(cfg)
    try:
        step = TR.TrainStep(m, opt, cfg.temperature, torch.device("cpu"), use_graph=True)
        fake_cuda.step = step
        a, b = make_batch(cfg, 2, 2, n_trans=4, seed=1), make_batch(cfg, 2, 2, n_trans=4, seed=2)
        step.prefetch(a)
        la = tuple(x.item() for x in step(None))
        step.prefetch(b)
        lb = tuple(x.item() for x in step(None))
        assert la != lb and len(step._graphs) == 1
        with pytest.raises(ValueError):
            TR.TrainStep(m, None, cfg.temperature, torch.device("cpu"), use_graph=True)(None)
    finally:
        opt.flat.release()
