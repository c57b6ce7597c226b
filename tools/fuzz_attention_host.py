"""Randomised sweep of the attention kernels on the CPU SIMT stand-in (tests/host_kernels): random batch / heads / mode / frames / group
sizes / sequence lengths / causal / key padding / head dim, forward + backward against the torch restatement of tests/emu.py.
Development tool (imports test infrastructure).   usage: python tools/fuzz_attention_host.py [generic|specialised] [seed] [seconds]"""
import ctypes
import os
import random
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import emu  # noqa: E402

HK = os.path.join(ROOT, "tests", "host_kernels")
BF16 = torch.bfloat16
P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else None)   # noqa: E731
I = lambda v: ctypes.c_longlong(int(v))                                  # noqa: E731


def build(src):
    out = os.path.join(tempfile.mkdtemp(prefix="hostk"), "lib.so")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-DTVTS_HOST_SHIM",
                    "-I", HK, os.path.join(HK, src), "-o", out], check=True)
    return ctypes.CDLL(out)


def check(lib, generic, B, H, mode, T, n, N, causal, d, klen, group):
    qkv = torch.randn(B, N, 3 * H * d).to(BF16)
    dout = torch.randn(B * N, H * d).to(BF16)
    scale = d ** -0.5
    out = torch.full((B * N, H * d), float("nan"), dtype=BF16)
    lse = torch.full((B, H, N), float("nan"))
    kinds = ctypes.c_int(0)
    if generic:
        used = lib.h_attn_fwd(P(qkv), P(out), P(lse), P(klen), I(B), I(N), I(H), I(d), I(mode), I(T), I(n), I(causal), ctypes.c_float(scale), ctypes.c_int(group))
    else:
        lib.h64_attn_fwd(P(qkv), P(out), P(lse), I(B), I(N), I(H), I(mode), I(T), I(n), I(causal), ctypes.c_float(scale), ctypes.byref(kinds))
        used = kinds.value
    ro, rl = torch.empty_like(out), torch.empty_like(lse)
    if klen is None:
        emu.attn_fwd(qkv, ro, rl, B, N, H, d, mode, T, n, causal, scale)
    else:
        emu.attn_padded_fwd(qkv, ro, rl, klen, B, N, H, d, scale)
    ok = torch.allclose(out.float(), ro.float(), atol=2e-2) and torch.allclose(lse, rl, atol=1e-4)
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty_like(lse)
    if generic:
        lib.h_attn_bwd(P(qkv), P(ro), P(dout), P(rl), P(delta), P(dqkv), P(klen), I(B), I(N), I(H), I(d), I(mode), I(T), I(n), I(causal),
                       ctypes.c_float(scale), ctypes.c_int(group))
    else:
        lib.h64_attn_bwd(P(qkv), P(ro), P(dout), P(rl), P(delta), P(dqkv), I(B), I(N), I(H), I(mode), I(T), I(n), I(causal), ctypes.c_float(scale))
    rd, rdel = torch.empty_like(qkv), torch.empty_like(lse)
    if klen is None:
        emu.attn_bwd(qkv, ro, dout, rl, rdel, rd, B, N, H, d, mode, T, n, causal, scale)
    else:
        emu.attn_padded_bwd(qkv, ro, dout, rl, rdel, rd, klen, B, N, H, d, scale)
    ok = ok and bool(torch.isfinite(dqkv.float()).all()) and torch.allclose(dqkv.float(), rd.float(), atol=3e-2, rtol=3e-2)
    return ok, used


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "generic"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 120.0
    generic = which == "generic"
    lib = build("harness_attn.cpp" if generic else "harness_attn64.cpp")
    random.seed(seed)
    torch.manual_seed(seed)
    t0, cases, kinds = time.time(), 0, {}
    while time.time() - t0 < seconds:
        d = random.choice([64, 80]) if generic else 64
        B, H, group = random.randint(1, 2), random.randint(1, 2), random.choice([0, 1, 1])
        mode = random.choice([0, 0, 1, 2])
        klen, causal, T, n = None, 0, 0, 0
        if mode == 0:
            N = random.choice([random.randint(1, 40), random.randint(41, 130), random.randint(131, 260)])
            r = random.random()
            if r < 0.3:
                causal = 1
            elif r < 0.55 and generic:
                klen = torch.tensor([random.randint(1, N) for _ in range(B)], dtype=torch.int32)
        elif mode == 1:
            T = random.randint(1, 4)
            n = random.choice([random.randint(1, 30), random.randint(31, 111), random.randint(112, 140)])
            N = 1 + T * n
        else:
            T = random.choice([random.randint(1, 15), random.randint(16, 31), random.randint(32, 40)])
            n = random.randint(1, 9)
            N = 1 + T * n
        ok, used = check(lib, generic, B, H, mode, T, n, N, causal, d, klen, group)
        cases += 1
        kinds[used] = kinds.get(used, 0) + 1
        if not ok:
            print("MISMATCH", dict(B=B, H=H, mode=mode, T=T, n=n, N=N, causal=causal, d=d, klen=None if klen is None else klen.tolist(), group=group))
            sys.exit(1)
    print(f"{which}: {cases} random cases, no mismatch; kernel selection histogram {kinds}; {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
