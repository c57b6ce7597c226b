"""GPU debugging aid: run a model step with every native call SHADOWED by the torch restatement (tests/emu.py) on cloned
arguments, and report, op by op, where the real kernel and the restatement disagree.   (test/debug infrastructure only)

  python tools/shadow_debug.py [cfg] [batch] [frames] [n_trans]
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import emu  # noqa: E402
from tvts_b200 import _lib as L  # noqa: E402
from tvts_b200 import config as C  # noqa: E402
from tvts_b200 import engine as E  # noqa: E402
from tvts_b200 import modules as M  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402

real_call, real_gemm = L.call, L.gemm
LOG = []


def _cmp(name, idx, a, b):
    if not a.is_floating_point():
        bad = (a != b).sum().item()
        if bad:
            LOG.append((name, idx, "int mismatch", bad))
            print(f"!! {name} arg{idx}: {bad} integer mismatches", flush=True)
        return
    a32, b32 = a.float(), b.float()
    fin = torch.isfinite(b32)
    if not torch.isfinite(a32[fin]).all():
        print(f"!! {name} arg{idx}: non-finite values in kernel output", flush=True)
    err = (a32 - b32)[fin].abs().max().item() if fin.any() else 0.0
    ref = b32[fin].abs().max().item() if fin.any() else 0.0
    tol = (2e-2 if a.dtype == torch.bfloat16 else 2e-3) * max(ref, 1.0)
    flag = "!!" if err > tol else "  "
    if err > tol or os.environ.get("SHADOW_VERBOSE"):
        print(f"{flag} {name} arg{idx} shape={tuple(a.shape)} dtype={a.dtype} max_err={err:.4g} ref_max={ref:.4g}", flush=True)
    if err > tol:
        LOG.append((name, idx, err, ref))


def shadow_call(name, *args):
    clones = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    real_call(name, *args)
    emu.OPS[name](*clones)
    torch.cuda.synchronize()
    desc = " ".join(str(a) for a in args if not isinstance(a, torch.Tensor) and a is not None)
    for i, (a, b) in enumerate(zip(args, clones)):
        if isinstance(a, torch.Tensor):
            _cmp(f"{name}({desc})", i, a, b)


def shadow_gemm(a, b, out, **kw):
    out2 = out.clone()
    kw2 = dict(kw)
    if kw.get("out_pre") is not None:
        kw2["out_pre"] = kw["out_pre"].clone()
    real_gemm(a, b, out, **kw)
    emu.gemm(a, b, out2, **kw2)
    torch.cuda.synchronize()
    desc = f"gemm(M={kw['M']} N={kw['N']} K={kw['K']} a_mn={kw.get('a_mn', 0)} b_mn={kw.get('b_mn', 0)} acc={kw.get('accumulate', 0)} act={kw.get('act')} dact={kw.get('dact')})"
    _cmp(desc, "out", out, out2)
    if kw.get("out_pre") is not None:
        _cmp(desc, "out_pre", kw["out_pre"], kw2["out_pre"])
    return out


def main():
    cfg = getattr(C, sys.argv[1]) if len(sys.argv) > 1 else C.TINY_B_MASK
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    n_trans = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    L.call, L.gemm = shadow_call, shadow_gemm
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    m = m.cuda()
    data = make_batch(cfg, batch, frames, n_trans=n_trans, seed=5)
    dev = {k: v.cuda() for k, v in data.items()}
    te, ve, pred = m(dev)
    loss1 = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te))
    loss2 = E.sort_ce(pred, dev["label"]) if pred is not None else 0.0
    (loss1 + loss2).backward()
    torch.cuda.synchronize()
    print(f"shadow run done: {len(LOG)} mismatching op outputs")
    for l in LOG[:40]:
        print("  ", l)


if __name__ == "__main__":
    main()
