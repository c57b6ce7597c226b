"""GPU debugging aid for timing-dependent faults: run the same forward(+backward) several times WITHOUT intermediate syncs,
recording a clone of every tensor argument after each native call, then report the first call whose recorded tensors differ
between runs or contain non-finite values.   (test/debug infrastructure only)

  python tools/replay_debug.py [cfg] [batch] [frames] [n_trans] [runs]
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from tvts_b200 import _lib as L  # noqa: E402
from tvts_b200 import config as C  # noqa: E402
from tvts_b200 import engine as E  # noqa: E402
from tvts_b200 import modules as M  # noqa: E402
from tvts_b200.synthetic import make_batch, make_state_dict  # noqa: E402

real_call, real_gemm = L.call, L.gemm
REC = []
RECORD = os.environ.get("REPLAY_NO_RECORD") is None


def rec_call(name, *args):
    real_call(name, *args)
    if RECORD:
        desc = name + "(" + " ".join(str(a) for a in args if not isinstance(a, torch.Tensor) and a is not None) + ")"
        REC.append((desc, [a.clone() if isinstance(a, torch.Tensor) else None for a in args]))


def rec_gemm(a, b, out, **kw):
    real_gemm(a, b, out, **kw)
    if RECORD:
        desc = f"gemm(M={kw['M']} N={kw['N']} K={kw['K']} a_mn={int(kw.get('a_mn', 0))} b_mn={int(kw.get('b_mn', 0))} acc={int(kw.get('accumulate', 0))} act={kw.get('act')} dact={kw.get('dact')})"
        REC.append((desc, [a.clone(), b.clone(), out.clone(), kw["out_pre"].clone() if kw.get("out_pre") is not None else None]))
    return out


def main():
    cfg = getattr(C, sys.argv[1]) if len(sys.argv) > 1 else C.TINY_B_MASK
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    n_trans = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    runs = int(sys.argv[5]) if len(sys.argv) > 5 else 6
    L.call, L.gemm = rec_call, rec_gemm
    m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
    m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
    m = m.cuda()
    data = make_batch(cfg, batch, frames, n_trans=n_trans, seed=5)
    dev = {k: (v.cuda() if k != "keep_ind" else v) for k, v in data.items()}
    ref = None
    for r in range(runs):
        REC.clear()
        for p in m.parameters():
            p.grad = None
        te, ve, pred = m(dev)
        loss1 = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te))
        loss2 = E.sort_ce(pred, dev["label"]) if pred is not None else 0.0
        (loss1 + loss2).backward()
        torch.cuda.synchronize()
        bad_pred = not torch.isfinite(pred).all().item()
        print(f"run {r}: loss1={loss1.item():.5f} loss2={float(loss2):.5f} pred finite={not bad_pred} calls={len(REC)}", flush=True)
        if not RECORD:
            continue
        cur = [(d, [None if t is None else t.cpu() for t in ts]) for d, ts in REC]
        first = None
        for i, (d, ts) in enumerate(cur):
            for j, t in enumerate(ts):
                if t is None or not t.is_floating_point():
                    continue
                if not torch.isfinite(t.float()).all():
                    first = (i, d, j, "non-finite")
                    break
                if ref is not None:
                    o = ref[i][1][j]
                    if not torch.equal(t, o) and "acc=1" not in d and "bwd" not in d and "colsum" not in d:
                        err = (t.float() - o.float()).abs().max().item()
                        if err > 1e-3 * max(1.0, o.float().abs().max().item()):
                            first = (i, d, j, f"differs from run 0 by {err:.4g}")
                            break
            if first:
                break
        if first:
            print(f"   first suspicious call #{first[0]}: {first[1]} arg{first[2]}: {first[3]}")
            lo = max(0, first[0] - 3)
            for k in range(lo, first[0] + 1):
                print("      ", k, cur[k][0])
        if ref is None:
            ref = cur


if __name__ == "__main__":
    main()
