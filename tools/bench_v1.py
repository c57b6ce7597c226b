"""Throughput of the TVTS v1 step (BASELINE.json configs[4]) on one GPU: ViT-B/16 with 2-frame tubelets over 16 frames, 49 of 196
patches kept per tube (N = 393 tokens), DistilBERT-base text encoder on 4 captions of 50 (padded) tokens per clip, projection heads,
sort head, both losses, backward, AdamW (one group, lr 1e-4, wd 0: v1/configs/dist-yt-pt.json:44-54).  Random-init weights, synthetic
inputs.  Written after the round-1 GPU budget was spent: NOT yet run on a GPU (bench.py stays the contract bench; this tool is the
starting point for a `c5` workload there).

    python tools/bench_v1.py [--batch 24] [--steps 10] [--warmup 3] [--no-graph]
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tvts_b200 import _lib, modules_v1 as V1, optim  # noqa: E402
from tvts_b200.trainer import TrainStep  # noqa: E402


def flops_per_pair(T=16, P=196, n=49, D=768, L=12, p=16, W=768, Lt=6, ctx=50, n_trans=4, E=768):
    """SURVEY.md section 8d row c5 (2*MAC, fwd+bwd = 3x fwd): 76.2 GF video + n_trans x 4.3 GF text + 12.2 GF sort head per pair forward."""
    nt = T // 2
    N = 1 + nt * n
    f_video = 2 * nt * P * (3 * 2 * p * p) * D + L * (24 * N * D * D + 4 * N * N * D)
    f_text = Lt * (24 * ctx * W * W + 4 * ctx * ctx * W)
    S = N + n_trans
    f_sort = 2 * (24 * S * E * E + 4 * S * S * E)
    return 3.0 * (f_video + n_trans * f_text + f_sort)


def make_batch(B, T=16, P=196, n=49, n_trans=4, ctx=50, vocab=30522, seed=0):
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 3, 224, 224, generator=g)
    keep = torch.stack([torch.stack([torch.randperm(P, generator=g)[:n] for _ in range(T // 2)]) for _ in range(B)])
    ids = torch.randint(1000, vocab, (n_trans * B, ctx), generator=g)
    lens = torch.randint(6, ctx + 1, (n_trans * B,), generator=g)
    mask = (torch.arange(ctx)[None, :] < lens[:, None]).long()
    return {"video": video, "keep_ind": keep, "label": torch.arange(n_trans).repeat(B, 1),
            "text": {"input_ids": ids * mask, "attention_mask": mask}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=24)          # v1/configs/dist-yt-pt.json:28 (per GPU)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise RuntimeError("tools/bench_v1.py: no CUDA device; the hot path has no CPU fallback")
    _lib.lib()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = V1.TVTS(types.SimpleNamespace(local_rank=0), {"num_frames": 16}, {"model": "distilbert-base-uncased", "pretrained": True},
                    text_model=V1.DistilBertShell()).to(dev)
    opt = optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0)
    step = TrainStep(model, opt, 0.05, dev, use_graph=not args.no_graph)
    data = make_batch(args.batch)
    resident = {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()}) for k, v in data.items()}
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        l1, l2 = step(resident)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    fl = flops_per_pair()
    print(json.dumps({"metric": "video-text pairs/sec (TVTS v1 ViT-B/16 tubelets, 16x224^2 frames) fwd+bwd+AdamW", "value": args.batch / (ms / 1e3),
                      "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "loss": (l1 + l2).item(),
                      "gflop_per_pair": fl / 1e9, "model_tflops": args.batch * fl / (ms / 1e3) / 1e12,
                      "config": {"workload": f"c5: TVTS v1 base_patch16_224, T=16, 49/196 patches per tube, DistilBERT-base x4 captions of 50 tokens, batch {args.batch}"}}))


if __name__ == "__main__":
    main()
