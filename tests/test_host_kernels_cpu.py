"""The streaming kernels written after the round-1 GPU budget was spent (tvts_b200/csrc/v1_glue.cu, input_stage.cu) EXECUTED ON THE
CPU: their bodies are compiled unmodified by g++ against tests/host_kernels/host_shim.h and run once per (block, thread) of the launch
grid (tests/host_kernels/harness.cpp), then compared with the torch restatements of tests/emu.py.  This checks the index arithmetic,
the bf16 rounding points and the zero padding without a GPU; the launch configuration and the device code generation remain for
the GPU tests (tests/test_zz_round1_unverified_gpu.py)."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch

import emu
from tvts_b200._lib import OPERAND

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels")
pytestmark = [pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++"),
              pytest.mark.skipif(OPERAND != "bf16", reason="the host shim models the bfloat16 build")]
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def hk(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostk") / "libhostk.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DTVTS_HOST_SHIM", "-I", HERE, os.path.join(HERE, "harness.cpp"), "-o", out],
                   check=True)
    return ctypes.CDLL(out)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def I(v):
    return ctypes.c_longlong(int(v))


def keep_per_tube(B, nt, Pn, n, g):
    return torch.stack([torch.stack([torch.randperm(Pn, generator=g)[:n] for _ in range(nt)]) for _ in range(B)]).contiguous()


@pytest.mark.parametrize("B,T,R,p,n", [(2, 4, 64, 16, 5), (1, 8, 96, 32, 9), (2, 2, 32, 8, 16)])
def test_tubelet_gather_and_assembly(hk, B, T, R, p, n):
    g = torch.Generator().manual_seed(R + T)
    nt, Pn, D = T // 2, (R // p) ** 2, 64
    video = torch.randn(B, T, 3, R, R, generator=g)
    keep = keep_per_tube(B, nt, Pn, n, g)
    got = torch.full((B * nt * n, 6 * p * p), 7.0, dtype=BF16)
    ref = torch.empty_like(got)
    hk.h_tubelet_gather(P(video), P(keep), P(got), I(B), I(T), I(R), I(p), I(n))
    emu.tubelet_gather(video, keep, ref, B, T, R, p, n)
    assert torch.equal(got, ref)
    tok, cls, pos, tem = torch.randn(B * nt * n, D, generator=g), torch.randn(1, 1, D, generator=g), torch.randn(1, Pn + 1, D, generator=g), \
        torch.randn(1, nt + 2, D, generator=g)
    x_got, x_ref = torch.full((B * (1 + nt * n), D), 7.0), torch.empty(B * (1 + nt * n), D)
    hk.h_assemble_tube(P(tok), P(cls), P(pos), P(tem), P(keep), P(x_got), I(B), I(nt), I(n), I(D))
    emu.video_assemble_tube(tok, cls, pos, tem, keep, x_ref, B, nt, n, D)
    assert torch.allclose(x_got, x_ref, atol=1e-6)
    dx0 = torch.randn(B * (1 + nt * n), D, generator=g)
    outs = []
    for fn in ("host", "emu"):
        dcls, dpos, dtem = torch.zeros(1, 1, D), torch.zeros(1, Pn + 1, D), torch.zeros(1, nt + 2, D)
        dtok = torch.full((B * nt * n, D), 7.0, dtype=BF16)
        if fn == "host":
            hk.h_assemble_tube_bwd(P(dx0), P(keep), P(dcls), P(dpos), P(dtem), P(dtok), I(B), I(nt), I(n), I(D))
        else:
            emu.video_assemble_tube_bwd(dx0, keep, dcls, dpos, dtem, dtok, B, nt, n, D)
        outs.append((dcls, dpos, dtem, dtok))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert torch.allclose(a, b, atol=1e-4)
    assert torch.equal(outs[0][3], outs[1][3])


def test_relu_kernels(hk):
    g = torch.Generator().manual_seed(1)
    x, dy = torch.randn(6, 40, generator=g), torch.randn(6, 40, generator=g)
    y_got, y_ref = torch.empty(6, 40, dtype=BF16), torch.empty(6, 40, dtype=BF16)
    hk.h_relu_bf16(P(x), P(y_got), I(240))
    emu.relu_bf16(x, y_ref, 240)
    assert torch.equal(y_got, y_ref)
    d_got, d_ref = torch.empty(6, 40), torch.empty(6, 40)
    hk.h_relu_bwd(P(x), P(dy), P(d_got), I(240))
    emu.relu_bwd(x, dy, d_ref, 240)
    assert torch.equal(d_got, d_ref)


@pytest.mark.parametrize("B,T,R,p,n", [(2, 3, 64, 16, 7), (1, 2, 64, 32, 4), (2, 1, 32, 8, 16)])
def test_fused_uint8_input_stage(hk, B, T, R, p, n):
    g = torch.Generator().manual_seed(R + p)
    Pn = (R // p) ** 2
    keep = torch.stack([torch.randperm(Pn, generator=g)[:n] for _ in range(B)]).contiguous()
    u8 = torch.randint(0, 256, (B, T, 3, R, R), dtype=torch.uint8, generator=g)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    got = torch.full((B * T * n, 3 * p * p), 7.0, dtype=BF16)
    ref = torch.empty_like(got)
    hk.h_patch_gather_u8(P(u8), P(keep), P(got), I(B), I(T), I(R), I(p), I(n), (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std))
    emu.patch_gather_u8(u8, keep, ref, B, T, R, p, n, mean, std)
    assert torch.equal(got, ref)          # x/255, (x-mean)/std in fp32 in the reference's order, then bf16: bit-identical


@pytest.mark.parametrize("B,T,R,p,n", [(2, 3, 56, 14, 5), (1, 2, 28, 14, 4), (1, 1, 36, 6, 9)])
def test_padded_patch_gather_and_weight_cast(hk, B, T, R, p, n):
    g = torch.Generator().manual_seed(p + n)
    Pn, K = (R // p) ** 2, 3 * p * p
    Kp = (K + 7) // 8 * 8
    keep = torch.stack([torch.randperm(Pn, generator=g)[:n] for _ in range(B)]).contiguous()
    video = torch.randn(B, T, 3, R, R, generator=g)
    got = torch.full((B * T * n, Kp), 7.0, dtype=BF16)
    ref = torch.full((B * T * n, Kp), 7.0, dtype=BF16)
    hk.h_patch_gather_ld(P(video), P(keep), P(got), I(B), I(T), I(R), I(p), I(n), I(Kp))
    emu.patch_gather_ld(video, keep, ref, B, T, R, p, n, Kp)
    assert torch.equal(got, ref)
    w = torch.randn(10, 3, p, p, generator=g)
    got = torch.full((10, Kp), 7.0, dtype=BF16)
    ref = torch.full((10, Kp), 7.0, dtype=BF16)
    hk.h_cast_pad(P(w), P(got), I(10), I(K), I(Kp))
    emu.cast_bf16_pad(w, ref, 10, K, Kp)
    assert torch.equal(got, ref)
