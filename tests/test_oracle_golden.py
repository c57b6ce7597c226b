"""The CPU oracle (oracle/tvts_oracle.py) must reproduce the outputs of the UNMODIFIED reference recorded in
tests/golden/*.npz by oracle/make_golden.py (the reference ships no fixtures of its own for this path)."""
import os

import numpy as np
import pytest
import torch

import tvts_oracle as O
from tvts_b200 import config as C
from tvts_b200.synthetic import make_batch, make_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"tiny_B": C.TINY_B, "tiny_B_mask": C.TINY_B_MASK, "tiny_B_cap": C.TINY_B, "c1_b32": C.TVTSV2_B_32, "tiny_H": C.TINY_H}


@pytest.mark.parametrize("name", ["tiny_B", "tiny_B_mask", "tiny_B_cap", "c1_b32", "tiny_H"])
def test_oracle_matches_reference_golden(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    torch.set_num_threads(os.cpu_count())
    sd = make_state_dict(cfg, seed=1234)
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    l1, l2, (te, ve, pr), grads = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    assert abs(l1.item() - float(g["loss1"])) < 2e-5
    assert abs(l2.item() - float(g["loss2"])) < 2e-5
    np.testing.assert_allclose(te.numpy(), g["text_emb"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(ve.numpy(), g["video_emb"], atol=2e-5, rtol=1e-4)
    if "pred_order" in g.files:
        np.testing.assert_allclose(pr.numpy(), g["pred_order"], atol=5e-5, rtol=1e-4)
    else:
        assert pr is None
    names = [str(s) for s in g["grad_names"]]
    assert set(names) == set(grads.keys())
    for k, nrm, head in zip(names, g["grad_norms"], g["grad_heads"]):
        gn = grads[k].double().norm().item()
        assert abs(gn - nrm) <= 1e-3 * nrm + 1e-7, (k, gn, nrm)
        h = grads[k].reshape(-1)[:8].double().numpy()
        np.testing.assert_allclose(h, head[: h.size], atol=1e-5 + 1e-3 * np.abs(head[: h.size]).max())
