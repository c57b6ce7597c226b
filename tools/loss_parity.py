"""Loss-trajectory parity: N optimizer steps of the B200 path (TrainStep: fwd + bwd + fused AdamW, CUDA-graph replay) against the
CPU oracle (reference algorithm in fp32 + the restated transformers.AdamW) from the same initial weights on the same synthetic
batches.  Prints the per-step losses of both and the maximum deviation.   usage: python tools/loss_parity.py [steps] [workload]

workloads: c1 = BASELINE.json configs[0] (TVTSv2 ViT-B/32, 2 frames, 4 pairs, n_trans 4);  c3s = 2 pairs of the headline shape
(ViT-B/16, mask 0.5, 8 frames);  tiny = the test-suite toy model;
tiny_h = the H/14-shaped toy model (width 640, head dim 80, 14x14 patches).   TVTS_OPERAND=fp16 selects the IEEE-half build."""
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import tvts_oracle as O  # noqa: E402
from tvts_b200 import config as C  # noqa: E402
from tvts_b200 import modules as M  # noqa: E402
from tvts_b200 import optim  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402
from tvts_b200.trainer import TrainStep  # noqa: E402


def oracle_run(cfg, sd0, groups_by_name, batches, frozen):
    sd = {k: v.clone() for k, v in sd0.items()}
    mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()}
    trainable = {k for k in sd if k not in frozen}
    out = []
    for t, data in enumerate(batches, 1):
        l1, l2, _, grads = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg, trainable=trainable)
        out.append((l1.item(), l2.item()))
        for k, g in grads.items():
            lr, wd = groups_by_name[k]
            O.adamw_step(sd[k], g, mom[k][0], mom[k][1], t, lr, weight_decay=wd)
    return out


def run(steps=100, workload="c1", lr_scale=1.0, use_graph=True, verbose=True):
    if workload == "c1":
        cfg, batch, frames = C.TVTSV2_B_32, 4, 2
    elif workload == "c3s":            # a 2-pair sample of the headline shape (ViT-B/16, mask 0.5, T = 8): what bench.py times
        cfg, batch, frames = C.TVTSV2_B_16, 2, 8
    elif workload == "tiny_h":
        cfg, batch, frames = C.TINY_H640, 2, 3
    else:
        cfg, batch, frames = C.TINY_B_MASK, 2, 3
    torch.set_num_threads(os.cpu_count() or 1)
    sd0 = make_state_dict(cfg, seed=1234)
    batches = [make_batch(cfg, batch, frames, n_trans=4, seed=100 + i) for i in range(steps)]
    m = (M.TVTSv2_H_14 if cfg.post_mode == "h14" else M.TVTSv2Base)(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(sd0, strict=True)
    m = m.cuda()
    opt = optim.build_reference_optimizer(m)
    for g in opt.param_groups:
        g["lr"] *= lr_scale
    names = {id(p): n for n, p in m.named_parameters()}
    groups_by_name = {names[id(p)]: (g["lr"], g["weight_decay"]) for g in opt.param_groups for p in g["params"]}
    frozen = {n for n, p in m.named_parameters() if not p.requires_grad}
    step = TrainStep(m, opt, cfg.temperature, torch.device("cuda"), use_graph=use_graph)
    ours = []
    t0 = time.time()
    try:
        for data in batches:
            l1, l2 = step(data)
            ours.append((l1.item(), l2.item()))
    finally:
        opt.flat.release()
    t_gpu = time.time() - t0
    t0 = time.time()
    ref = oracle_run(cfg, sd0, groups_by_name, batches, frozen)
    t_cpu = time.time() - t0
    d1 = max(abs(a[0] - b[0]) for a, b in zip(ours, ref))
    d2 = max(abs(a[1] - b[1]) for a, b in zip(ours, ref))
    dt = max(abs(a[0] + a[1] - b[0] - b[1]) for a, b in zip(ours, ref))
    if verbose:
        print(f"# {workload}: {cfg.name} batch {batch} frames {frames} n_trans 4, {steps} steps, lr x{lr_scale}; B200 {t_gpu:.1f}s, CPU oracle {t_cpu:.1f}s")
        print("# step  loss1(ours) loss1(oracle)  loss2(ours) loss2(oracle)")
        for i, (a, b) in enumerate(zip(ours, ref)):
            if i < 10 or i % 10 == 9:
                print(f"{i + 1:5d}  {a[0]:10.5f} {b[0]:10.5f}   {a[1]:10.5f} {b[1]:10.5f}")
        print(f"# max |d loss1| = {d1:.2e}   max |d loss2| = {d2:.2e}   max |d total| = {dt:.2e}")
    return d1, d2, dt


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 100, sys.argv[2] if len(sys.argv) > 2 else "c1",
        float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
