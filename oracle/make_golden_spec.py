"""Deterministic weights for an arbitrary (name, shape) list, shared by oracle/make_golden.py (which needs the reference tree) and
the fixture tests (which must not).  TEST INFRASTRUCTURE."""
import torch


def spec_state_dict(names, shapes, seed):
    """N(0, 0.05) in sorted-name order; 1-D `*weight` tensors (LayerNorm gains) are 1 + 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for n, shp in sorted(zip(names, shapes)):
        t = torch.randn(*shp, generator=g, dtype=torch.float32) * 0.05
        if n.endswith("weight") and len(shp) == 1:
            t = 1.0 + 2.0 * t
        sd[n] = t
    return sd
