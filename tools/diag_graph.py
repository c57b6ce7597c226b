"""GPU diagnostic for the eager-vs-CUDA-graph divergence of TrainStep (round-1 VERDICT item 1).   (test/debug infrastructure only)

Part 1: lr = 0, ONE batch repeated: every step must reproduce the same gradient arena.  Eager and graph runs are compared step by step
        with themselves (a path that differs from itself has a race / reads uninitialised memory) and with each other.
Part 2: the failing test's recipe (lr x30, 4 different batches): per step, which parameters' gradients / weights differ first.

  python tools/diag_graph.py [cfg] [steps]         env DIAG_SIDE=0 turns the attention side stream off for every run
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from tvts_b200 import _lib as L  # noqa: E402
from tvts_b200 import config as C  # noqa: E402
from tvts_b200 import modules as M  # noqa: E402
from tvts_b200 import optim  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402
from tvts_b200.trainer import TrainStep  # noqa: E402


def run(cfg, use_graph, n_steps, lr_scale, same_batch, side=1, batch=2, frames=3):
    L.lib().tvts_attn_set_side_stream(int(side))
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    m = m.cuda()
    opt = optim.build_reference_optimizer(m)
    for g in opt.param_groups:
        g["lr"] *= lr_scale
    step = TrainStep(m, opt, cfg.temperature, torch.device("cuda"), use_graph=use_graph)
    names = {id(p): n for n, p in m.named_parameters()}
    fs = opt.flat
    rec = []
    try:
        for it in range(n_steps):
            data = make_batch(cfg, batch, frames, n_trans=4, seed=0 if same_batch else it)
            l1, l2 = step(data)
            torch.cuda.synchronize()
            rec.append(dict(l1=l1.item(), l2=l2.item(), g=fs.g.clone().cpu(), p=fs.p.clone().cpu()))
        layout = [(names[id(p)], fs.offsets[i], p.numel()) for i, p in enumerate(fs.params)]
    finally:
        fs.release()
    return rec, layout


def diff(tag, a, b, layout, key, top=6):
    rows = []
    for name, off, n in layout:
        x, y = a[key][off:off + n], b[key][off:off + n]
        if not torch.equal(x, y):
            d = (x - y).abs().max().item()
            rows.append((d / (y.abs().max().item() + 1e-30), d, name))
    if not rows:
        print(f"    {tag}: {key} bit-identical")
        return False
    rows.sort(reverse=True)
    print(f"    {tag}: {key} differs in {len(rows)}/{len(layout)} tensors; worst (rel-to-max, abs, name):")
    for r in rows[:top]:
        print(f"        {r[0]:.3e} {r[1]:.3e} {r[2]}")
    return True


def main():
    cfg = getattr(C, sys.argv[1]) if len(sys.argv) > 1 else C.TINY_B_MASK
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    side = int(os.environ.get("DIAG_SIDE", "1"))
    print(f"== part 1: lr=0, same batch, {steps} steps, side stream {side}")
    runs = {}
    for tag, g in (("eager", False), ("graph", True)):
        rec, layout = run(cfg, g, steps, 0.0, True, side)
        runs[tag] = rec
        print(f"  {tag}: losses " + " ".join(f"{r['l1']:.7f}/{r['l2']:.7f}" for r in rec))
        for i in range(1, steps):
            diff(f"{tag} step{i} vs step0", rec[i], rec[0], layout, "g")
    for i in range(steps):
        diff(f"graph vs eager step{i}", runs["graph"][i], runs["eager"][i], layout, "g")
    if os.environ.get("DIAG_PART") == "1":
        return
    print("== part 2: lr x30, different batches (the failing test's recipe)")
    for s in ((1, 0) if side else (0,)):
        ea, layout = run(cfg, False, 5, 30.0, False, s)
        eb, _ = run(cfg, False, 5, 30.0, False, s)
        ga, _ = run(cfg, True, 5, 30.0, False, s)
        gb, _ = run(cfg, True, 5, 30.0, False, s)
        print(f"  side={s}")
        for tag, r in (("eagerA", ea), ("eagerB", eb), ("graphA", ga), ("graphB", gb)):
            print(f"   {tag}: " + " ".join(f"{x['l1']:.7f}/{x['l2']:.7f}" for x in r))
        for i in range(5):
            print(f"   step {i}")
            diff("eagerB vs eagerA", eb[i], ea[i], layout, "g")
            diff("graphB vs graphA", gb[i], ga[i], layout, "g")
            bad = diff("graphA vs eagerA", ga[i], ea[i], layout, "g")
            bad |= diff("graphA vs eagerA", ga[i], ea[i], layout, "p")
            if bad:
                break


if __name__ == "__main__":
    main()
