"""Which operand format does the loss-parity target need?  Runs the forward pass of the host path on the torch emulation of the
kernels (tests/emu.py: rounds to the operand dtype at every point the kernels do) in the chosen format, against the oracle.
Development tool (imports test infrastructure); results are quoted in DESIGN.md section 2.

    python tools/operand_format_experiment.py bf16|fp16|fp32
"""
import os
import sys
import types

import torch

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
if mode == "fp16":
    os.environ["TVTS_OPERAND"] = "fp16"          # tvts_b200._lib.OPERAND_DTYPE -> engine / emu operand dtype
elif mode == "fp32":
    os.environ["TVTS_OPERAND"] = "bf16"
    torch.bfloat16 = torch.float32               # experiment only: no 16-bit rounding anywhere (all deviation is operand rounding)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import emu  # noqa: E402
import tvts_oracle as O  # noqa: E402
from tvts_b200 import config as C, engine as E, modules as M  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402

emu.install()
if mode == "fp32":                               # the bf16-rows-as-fp32-words trick of these two helpers does not apply
    E.gather_rows_any = lambda src, idx, rows: src[idx].clone()

    def _scatter(src, idx, total_rows):
        dst = torch.zeros((total_rows, src.shape[1]), dtype=src.dtype)
        dst[idx] = src
        return dst
    E.scatter_rows_into_zeros = _scatter


def run(cfg, B, T, seeds):
    out = []
    sd = make_state_dict(cfg, seed=1234)
    for s in seeds:
        E.WEIGHTS.clear()
        m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
        m.load_state_dict(sd, strict=True)
        data = make_batch(cfg, B, T, n_trans=4, seed=s)
        with torch.no_grad():
            te, ve, pred = m(data)
            l1 = M.NormSoftmaxLoss(0.05)(M.sim_matrix(ve, te)).item()
            l2 = E.sort_ce(pred, data["label"], 2.0).item()
            ote, ove, opred = O.model_forward(sd, data["text"], data["video"], data["keep_ind"], cfg)
            o1 = O.norm_softmax_loss(O.sim_matrix(ove, ote), 0.05).item()
            o2 = O.sort_ce(opred, data["label"]).item()
        out.append((abs(l1 - o1), abs(l2 - o2), (te - ote).abs().max().item()))
    return out


for cfg, B, T, name in ((C.TINY_B_MASK, 2, 3, "tiny"), (C.TVTSV2_B_32, 4, 2, "c1")):
    r = run(cfg, B, T, [100, 101, 102] if name == "c1" else list(range(100, 108)))
    print(mode, name, "max d1 %.2e max d2 %.2e  mean d1 %.2e d2 %.2e  max|d text_emb| %.2e" % (
        max(a for a, _, _ in r), max(b for _, b, _ in r), sum(a for a, _, _ in r) / len(r), sum(b for _, b, _ in r) / len(r),
        max(c for _, _, c in r)))
