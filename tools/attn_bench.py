"""CUDA-event timing of the tcgen05 attention kernels at the c3 shape (B=32, H=12, T=8, n=98), L2 flushed between launches.
usage: [TVTS_LIB_PATH=...] python tools/attn_bench.py      (measurement tool)"""
import numpy as np
import torch

from tvts_b200 import _lib as L

B, H, T, n, d = 32, 12, 8, 98, 64
N = 1 + T * n
dev = torch.device("cuda")
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * d, device=dev).to(L.OPERAND_DTYPE)
dout = torch.randn(B * N, H * d, device=dev).to(L.OPERAND_DTYPE)
out = torch.empty(B * N, H * d, device=dev, dtype=L.OPERAND_DTYPE)
lse = torch.empty(B, H, N, device=dev)
dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * H * d, device=dev)
delta = torch.empty_like(lse)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
scale = d ** -0.5


def timed(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = np.array(ts[3:])
    return ts.mean(), ts.min()


for mode in (1, 2):
    f = lambda: L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, 0, scale)
    b = lambda: L.call("attn_bwd_bias", qkv, out, dout, lse, delta, dqkv, dbias, B, N, H, d, mode, T, n, 0, scale)
    f()
    mf, lf = timed(f)
    mb, lb = timed(b)
    print(f"mode {mode}: fwd {mf:.1f} us (min {lf:.1f})   bwd {mb:.1f} us (min {lb:.1f})   [HBM floor: fwd 23.5, bwd 41 us]")
