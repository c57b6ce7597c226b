"""Phase timing of the persistent tcgen05 attention BACKWARD kernel (measurement tool; needs a library built with -DTVTS_ATTN_PROF, see the build line in
tools/r2_call15.sh):  TVTS_LIB_PATH=build_ab/prof_fp16.so python tools/attn_phase_prof.py [mode]
Thread 0 of every CTA stamps clock64 at the phase boundaries; this prints the mean cycles per phase over the tiles of one launch at the
c3 shape (B=32, H=12, T=8, n=98) and the gap between consecutive CTAs of one SM slot."""
import ctypes
import sys

import numpy as np
import torch

from tvts_b200 import _lib as L

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, H, T, n, d = 32, 12, 8, 98, 64
N = 1 + T * n
dev = torch.device("cuda")
torch.manual_seed(0)
qkv = (torch.randn(B, N, 3 * H * d, device=dev)).to(L.OPERAND_DTYPE)
dout = torch.randn(B * N, H * d, device=dev).to(L.OPERAND_DTYPE)
out = torch.empty(B * N, H * d, device=dev, dtype=L.OPERAND_DTYPE)
lse = torch.empty(B, H, N, device=dev)
dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * H * d, device=dev)
scale = d ** -0.5
buf = np.zeros((8192, 13), dtype=np.int64)


def stamps():
    torch.cuda.synchronize()
    assert L.lib().tvts_attn_tc_prof_read(ctypes.c_void_p(buf.ctypes.data)) == 0
    return buf.copy()


def report(name, st, ntiles, labels):
    t = st[:ntiles, :len(labels) + 1].astype(np.float64)
    print(f"== {name}: {ntiles} tiles; mean cycles per phase of the persistent tile loop (thread 0)")
    for i, lab in enumerate(labels):
        dphase = t[:, i + 1] - t[:, i]
        print(f"  {lab:52s} {dphase.mean():8.0f}  (p10 {np.percentile(dphase, 10):6.0f}  p90 {np.percentile(dphase, 90):6.0f})")
    life = t[:, len(labels)] - t[:, 0]
    print(f"  {'one loop iteration':52s} {life.mean():8.0f}")
    sm = st[:ntiles, 12]
    print(f"  tiles per SM {ntiles / len(np.unique(sm)):.1f}")
    ns0, ns1, cta = st[:ntiles, 10], st[:ntiles, 9], st[:ntiles, 11]
    print(f"  kernel span by globaltimer: {(ns1.max() - ns0.min()) / 1e3:.1f} us; first tile start spread {(ns0[:int(cta.max()) + 1].max() - ns0.min()) / 1e3:.1f} us")
    per = {}
    for c_, a_, b_, cyc in zip(cta, ns0, ns1, life):
        e = per.setdefault(int(c_), [a_, b_, 0, 0.0])
        e[0] = min(e[0], a_); e[1] = max(e[1], b_); e[2] += 1; e[3] += cyc
    busy = np.array([(v[1] - v[0]) / 1e3 for v in per.values()])
    ntile = np.array([v[2] for v in per.values()])
    cyc = np.array([v[3] for v in per.values()])
    print(f"  per CTA: tiles {ntile.min()}..{ntile.max()}, first-start-to-last-end {busy.mean():.1f} us (min {busy.min():.1f} max {busy.max():.1f}), "
          f"loop cycles {cyc.mean():.0f} (min {cyc.min():.0f} max {cyc.max():.0f}) -> implied clock {cyc.mean() / busy.mean() / 1e3:.2f} GHz")


L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, 0, scale)      # produces out / lse for the backward (the forward kernel carries no stamps)
chunks = T if mode == 1 else 7
for _ in range(3):
    L.call("attn_bwd_bias", qkv, out, dout, lse, torch.empty_like(lse), dqkv, dbias, B, N, H, d, mode, T, n, 0, scale)
st = stamps()
lab_b = ["wait TMA boxes + S, dP MMAs (bar_1; previous CLS partial published)", "pass A (P, delta), sync", "dV MMA (bar_dv; CLS ticket drawn)",
         "pass B (dS), sync", "dV epilogue + bias (under the dK, dQ MMAs)", "wait dK, dQ MMAs (bar_2)", "dQ, dK epilogue + bias",
         "next tile's setup, (merge), loop-end sync"]
report(f"backward mode {mode}", st, B * H * chunks, lab_b)
