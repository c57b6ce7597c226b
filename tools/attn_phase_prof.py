"""Phase timing of the tcgen05 attention kernels (measurement tool; needs a library built with -DTVTS_ATTN_PROF, see the build line in
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
    st = st[:ntiles]
    t = st[:, :len(labels) + 1].astype(np.float64)
    print(f"== {name}: {ntiles} tiles; mean cycles per phase (thread 0)")
    for i, lab in enumerate(labels):
        dphase = t[:, i + 1] - t[:, i]
        print(f"  {lab:38s} {dphase.mean():8.0f}  (p10 {np.percentile(dphase, 10):6.0f}  p90 {np.percentile(dphase, 90):6.0f})")
    life = t[:, len(labels)] - t[:, 0]
    print(f"  {'CTA lifetime':38s} {life.mean():8.0f}")
    # gaps between consecutive CTAs on one SM: sort by start per SM; with R resident CTAs per SM the k-th start follows the (k-R)-th end
    sm = st[:, 12]
    gaps = []
    for s_ in np.unique(sm):
        idx = np.where(sm == s_)[0]
        starts = np.sort(t[idx, 0])
        ends = np.sort(t[idx, len(labels)])
        R = int((starts < ends[0]).sum())
        for k in range(R, len(starts)):
            gaps.append(starts[k] - ends[k - R])
    if gaps:
        print(f"  gap end(CTA k-R) -> start(CTA k), same SM    {np.mean(gaps):8.0f}  (p10 {np.percentile(gaps, 10):6.0f}  p90 {np.percentile(gaps, 90):6.0f}); resident per SM ~{R}")
    tiles_per_sm = ntiles / len(np.unique(sm))
    span = max(t[:, len(labels)].max() - t[:, 0].min(), 1)
    print(f"  tiles per SM {tiles_per_sm:.1f}")


for _ in range(3):
    L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, 0, scale)
chunks = T if mode == 1 else 7
st = stamps()
lab_f = ["setup (alloc, barriers, CLS rows) + sync", "TMA load + S MMA (bar_s)", "softmax -> P, sync", "PV MMA (bar_o)", "epilogue staging, sync",
         "store issue + CLS publish/merge", "bulk_wait_read"]
stf = st[:, [0, 7, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12]]
report(f"forward mode {mode}", stf, B * H * chunks, lab_f)
for _ in range(3):
    L.call("attn_bwd_bias", qkv, out, dout, lse, torch.empty_like(lse), dqkv, dbias, B, N, H, d, mode, T, n, 0, scale)
st = stamps()
lab_b = ["setup (alloc, CLS rows, lse, delta) + sync", "TMA load + S, dP MMAs (bar_1)", "pass A (P, delta), sync", "dV MMA (bar_dv)", "pass B (dS), sync",
         "dV epilogue (under dK, dQ MMAs)", "wait dK, dQ MMAs (bar_2)", "dQ, dK epilogue, sync", "store issue, bias sums, CLS publish", "bulk_wait_read"]
report(f"backward mode {mode}", st, B * H * chunks, lab_b)
