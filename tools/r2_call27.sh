#!/bin/bash
# round 2, GPU call 27 (1 GPU): compute-sanitizer memcheck over the attention and GEMM kernel tests of the final code (register stores must
# never leave the matrix), then a repeat-run stress of the persistent attention backward (scheduler / ticket re-arming across launches)
set -x
O=gpurun_out/r2c27
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention and not query_window" -p no:cacheprovider > $O/memcheck_attention.log 2>&1; echo "memcheck attention rc=$?" | tee $O/rc.txt; tail -4 $O/memcheck_attention.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm" -p no:cacheprovider > $O/memcheck_gemm.log 2>&1; echo "memcheck gemm rc=$?" | tee -a $O/rc.txt; tail -4 $O/memcheck_gemm.log
PYTHONPATH=. timeout 600 python - > $O/stress.log 2>&1 <<'PY'
import torch
from tvts_b200 import _lib as L
torch.manual_seed(0)
dev = "cuda"
d = 64
ok = True
for (B, H, mode, T, n) in [(32, 12, 1, 8, 98), (32, 12, 2, 8, 98), (5, 3, 1, 8, 98), (7, 12, 2, 8, 23), (128, 8, 0, 0, 0)]:
    N = 77 if mode == 0 else 1 + T * n
    qkv = torch.randn(B, N, 3 * H * d, device=dev).to(L.OPERAND_DTYPE)
    dout = torch.randn(B * N, H * d, device=dev).to(L.OPERAND_DTYPE)
    out = torch.empty(B * N, H * d, device=dev, dtype=L.OPERAND_DTYPE)
    lse = torch.empty(B, H, N, device=dev)
    outs, grads, biases = [], [], []
    for rep in range(40):
        out.fill_(float("nan")); lse.fill_(float("nan"))
        L.call("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, int(mode == 0), d ** -0.5)
        dqkv = torch.full_like(qkv, float("nan"))
        dbias = torch.zeros(3 * H * d, device=dev)
        L.call("attn_bwd_bias", qkv, out, dout, lse, torch.empty_like(lse), dqkv, dbias, B, N, H, d, mode, T, n, int(mode == 0), d ** -0.5)
        if rep == 0:
            o0, l0, g0, b0 = out.clone(), lse.clone(), dqkv.clone(), dbias.clone()
        else:
            same = torch.equal(out, o0) and torch.equal(lse, l0) and torch.equal(dqkv, g0)
            fin = bool(torch.isfinite(dqkv.float()).all()) and bool(torch.isfinite(out.float()).all())
            bdiff = (dbias - b0).abs().max().item() / max(b0.abs().max().item(), 1e-6)
            if not (same and fin and bdiff < 1e-4):
                ok = False
                print("MISMATCH", (B, H, mode, T, n), "rep", rep, same, fin, bdiff)
                break
    print("shape", (B, H, mode, T, n), "40 repeats identical:", ok)
print("STRESS", "OK" if ok else "FAILED")
PY
tail -7 $O/stress.log
