"""Inputs / weights of the tests/golden/tiny_v1_full.npz case, regenerated from the seeds the fixture stores (test helper)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_spec import spec_state_dict  # noqa: E402


def load(name="tiny_v1_full"):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    D, heads, depth, patch, res, frames, nt, B, proj, Lc, vocab = [int(v) for v in g["dims"]]
    names = [str(s) for s in g["names"]]
    shapes = [tuple(int(x) for x in str(s).split(",")) for s in g["shapes"]]
    sd = spec_state_dict(names, shapes, int(g["wseed"]))
    video = torch.randn(B, frames, 3, res, res, generator=torch.Generator().manual_seed(int(g["video_seed"])))
    text = {"input_ids": torch.from_numpy(g["input_ids"]), "attention_mask": torch.from_numpy(g["attention_mask"])}
    data = {"text": text, "video": video, "keep_ind": torch.from_numpy(g["keep_ind"]), "label": torch.arange(nt).repeat(B, 1)}
    dims = types.SimpleNamespace(D=D, heads=heads, depth=depth, patch=patch, res=res, frames=frames, nt=nt, B=B, proj=proj, Lc=Lc, vocab=vocab)
    cfg = types.SimpleNamespace(patch=patch, width=D, heads=heads, layers=depth, sort_heads=heads, sort_depth=2, sort_ln_eps=1e-6)
    return g, dims, cfg, names, sd, data
