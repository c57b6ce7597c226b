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
    t = st[:ntiles, :len(labels) + 1].astype(np.float64)
    print(f"== {name}: {ntiles} tiles; mean cycles per phase of the persistent tile loop (thread 0)")
    for i, lab in enumerate(labels):
        dphase = t[:, i + 1] - t[:, i]
        print(f"  {lab:52s} {dphase.mean():8.0f}  (p10 {np.percentile(dphase, 10):6.0f}  p90 {np.percentile(dphase, 90):6.0f})")
    life = t[:, len(labels)] - t[:, 0]
    print(f"  {'one loop iteration':52s} {life.mean():8.0f}")
    sm = st[:ntiles, 12]
    print(f"  tiles per SM {ntiles / len(np.unique(sm)):.1f}")


for _ in range(3):
    L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, 0, scale)
chunks = T if mode == 1 else 7
st = stamps()
lab_f = ["wait TMA boxes + S MMA (bar_s)", "softmax -> P, sync", "P V MMA (bar_o; previous tile's CLS ticket drawn)", "epilogue staging, sync",
         "store issue, CLS publish, next tile's setup, (merge)", "store read-out wait + loop-end sync"]
report(f"forward mode {mode}", st, B * H * chunks, lab_f)
for _ in range(3):
    L.call("attn_bwd_bias", qkv, out, dout, lse, torch.empty_like(lse), dqkv, dbias, B, N, H, d, mode, T, n, 0, scale)
st = stamps()
lab_b = ["wait TMA boxes + S, dP MMAs (bar_1)", "pass A (P, delta), sync", "dV MMA (bar_dv; previous tile's CLS ticket drawn)", "pass B (dS), sync",
         "dV epilogue (under the dK, dQ MMAs)", "wait dK, dQ MMAs (bar_2)", "dQ, dK epilogue, sync",
         "stores, CLS publish, next tile's setup, bias sums, (merge)", "store read-out wait + loop-end sync"]
report(f"backward mode {mode}", st, B * H * chunks, lab_b)
