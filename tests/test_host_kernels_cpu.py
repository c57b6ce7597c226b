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


# ------------------------------------------------------------------------------------------------ tensor-core attention kernels
@pytest.fixture(scope="module")
def hattn(tmp_path_factory):
    """attention_hd.cu compiled for the CPU SIMT stand-in (tests/host_kernels/host_simt.h): CUDA threads = OS threads, ldmatrix / mma.sync /
    cp.async / shfl restated with their architectural layouts."""
    out = str(tmp_path_factory.mktemp("hostattn") / "libhostattn.so")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-DTVTS_HOST_SHIM", "-I", HERE,
                    os.path.join(HERE, "harness_attn.cpp"), "-o", out], check=True)
    return ctypes.CDLL(out)


ATTN_HOST_CASES = [
    # B, H, mode, T, n, N, causal, d, padded   (each case runs with the group-resident kernels allowed and with streamed kernels only)
    (2, 2, 0, 0, 0, 77, True, 80, False),      # causal, two streamed tiles
    (1, 2, 0, 0, 0, 150, False, 80, False),    # full, three tiles, ragged last tile
    (2, 2, 1, 2, 49, 99, False, 80, False),    # space
    (1, 2, 2, 3, 20, 61, False, 80, False),    # time (strided groups)
    (1, 2, 1, 3, 76, 229, False, 80, False),   # space at the H/14 kept-patch count: two stationary chunks per frame
    (1, 2, 2, 16, 6, 97, False, 80, False),    # time, 16 frames (c4)
    (2, 2, 1, 2, 49, 99, False, 64, False),    # the same code at head dim 64 (layouts shared with the GPU-verified kernels)
    (2, 1, 0, 0, 0, 130, True, 64, False),
    (5, 3, 0, 0, 0, 50, False, 64, True),      # key-padded (DistilBERT, v1)
    (2, 2, 0, 0, 0, 70, False, 80, True),
    (1, 1, 1, 16, 76, 1217, False, 80, False),  # the full c4 (H/14, 16 frames) token count, one head
    (2, 1, 0, 0, 0, 1, False, 80, False),       # a single token
    (1, 2, 0, 0, 0, 128, True, 80, False),      # sequence = exactly two tiles, causal
    (1, 2, 1, 2, 100, 201, False, 80, False),   # space, 101 rows per group: 7 warps
    (1, 2, 1, 2, 20, 41, False, 80, False),     # space, 21 rows: 2 warps
    (2, 2, 0, 0, 0, 40, False, 80, False),      # short full-attention sequence: group-resident in mode 0
    (1, 2, 0, 0, 0, 112, True, 64, False),      # largest group (7 warps), causal
    (1, 2, 2, 12, 76, 913, False, 80, False),   # time attention at the shipped H/14 clip length (3 frames x 4 clips): warp-per-slot kernels
    (1, 2, 2, 15, 6, 91, False, 80, False),     # the largest T the warp-per-slot kernels take
    (2, 1, 2, 8, 5, 41, False, 64, False),
    (1, 1, 2, 16, 76, 1217, False, 80, False),  # time attention of c4 (16 frames): two row tiles per slot
    (1, 2, 2, 31, 2, 63, False, 80, False),     # the largest T of the two-tile kernels
    (1, 2, 2, 32, 2, 65, False, 80, False),     # beyond it: streamed
    (2, 1, 2, 17, 5, 86, False, 64, False),
]


@pytest.mark.parametrize("group", [1, 0])
@pytest.mark.parametrize("B,H,mode,T,n,N,causal,d,padded", ATTN_HOST_CASES)
def test_generic_attention_kernels_on_the_cpu_simt_stand_in(hattn, B, H, mode, T, n, N, causal, d, padded, group):
    # kernel selection of hd_launch_fwd / hd_launch_bwd: 0 streamed, 1 group-resident (+ CLS launch), 2 / 3 warp-per-slot time kernels
    # with 1 / 2 row tiles per slot (+ CLS launch)
    fits = 0
    if not padded:
        if mode == 2:
            fits = 2 if T <= 15 else (3 if T <= 31 else 0)
        elif 2 <= ((N if mode == 0 else n + 1) + 15) // 16 <= 7:
            fits = 1
    if not group and not fits:
        pytest.skip("the streamed kernels are what runs in both settings")
    torch.manual_seed(N + mode + d)
    qkv = torch.randn(B, N, 3 * H * d).to(BF16)
    dout = torch.randn(B * N, H * d).to(BF16)
    scale = d ** -0.5
    klen = None
    if padded:
        klen = torch.randint(1, N + 1, (B,), dtype=torch.int32)
        klen[0] = N
        klen[-1] = 1                               # a sequence with a single valid key
    kp = ctypes.c_void_p(klen.data_ptr()) if padded else ctypes.c_void_p(None)
    out = torch.full((B * N, H * d), float("nan"), dtype=BF16)
    lse = torch.full((B, H, N), float("nan"))
    used = hattn.h_attn_fwd(P(qkv), P(out), P(lse), kp, I(B), I(N), I(H), I(d), I(mode), I(T), I(n), I(int(causal)), ctypes.c_float(scale),
                            ctypes.c_int(group))
    assert used == (fits if group else 0)
    ro, rl = torch.empty_like(out), torch.empty_like(lse)
    if padded:
        emu.attn_padded_fwd(qkv, ro, rl, klen, B, N, H, d, scale)
    else:
        emu.attn_fwd(qkv, ro, rl, B, N, H, d, mode, T, n, int(causal), scale)
    assert torch.allclose(out.float(), ro.float(), atol=2e-2), (out.float() - ro.float()).abs().max()
    assert torch.allclose(lse, rl, atol=1e-4)
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty_like(lse)
    assert hattn.h_attn_bwd(P(qkv), P(ro), P(dout), P(rl), P(delta), P(dqkv), kp, I(B), I(N), I(H), I(d), I(mode), I(T), I(n), I(int(causal)),
                            ctypes.c_float(scale), ctypes.c_int(group)) == used
    rd, rdel = torch.empty_like(qkv), torch.empty_like(lse)
    if padded:
        emu.attn_padded_bwd(qkv, ro, dout, rl, rdel, rd, klen, B, N, H, d, scale)
    else:
        emu.attn_bwd(qkv, ro, dout, rl, rdel, rd, B, N, H, d, mode, T, n, int(causal), scale)
    assert torch.isfinite(dqkv.float()).all(), "backward left elements unwritten"
    assert torch.allclose(dqkv.float(), rd.float(), atol=3e-2, rtol=3e-2), (dqkv.float() - rd.float()).abs().max()
    assert torch.allclose(delta, rdel, atol=1e-3)
    if padded:                                   # masked keys receive exactly zero dk / dv
        g5 = dqkv.view(B, N, 3, H, d)
        for b in range(B):
            assert (g5[b, int(klen[b]):, 1:] == 0).all()


# ------------------------------------------------------------------------------------------------ the specialised head-dim-64 kernels
@pytest.fixture(scope="module")
def hattn64(tmp_path_factory):
    """attention.cu (the GPU-verified streamed / group-resident / time / CLS kernels) on the same stand-in, with the real kernel selection."""
    out = str(tmp_path_factory.mktemp("hostattn64") / "libhostattn64.so")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-DTVTS_HOST_SHIM",
                    "-I", HERE, os.path.join(HERE, "harness_attn64.cpp"), "-o", out], check=True)
    return ctypes.CDLL(out)


STREAMED, GROUP, TIME, CLS = 1, 2, 4, 8
ATTN64_CASES = [
    # B, H, mode, T, n, N, causal, kernels the dispatch must pick
    (1, 2, 0, 0, 0, 77, True, GROUP),            # GPU-verified list: the stand-in must agree with what is known to be right on a B200
    (1, 1, 0, 0, 0, 200, False, STREAMED),
    (1, 2, 1, 2, 49, 99, False, GROUP | CLS),
    (1, 2, 2, 8, 5, 41, False, TIME | CLS),
    (1, 1, 2, 16, 6, 97, False, STREAMED),
    (2, 2, 0, 0, 0, 16, True, STREAMED),         # causal text sequences trimmed to the longest caption of a batch (trainer.trim_text_context):
    (2, 2, 0, 0, 0, 24, True, GROUP),            # group-resident kernels with 2..6 warps in causal mode, not part of the GPU test list
    (2, 2, 0, 0, 0, 40, True, GROUP),
    (2, 2, 0, 0, 0, 48, True, GROUP),
    (2, 2, 0, 0, 0, 64, True, GROUP),
    (2, 2, 0, 0, 0, 96, True, GROUP),
]


@pytest.mark.parametrize("B,H,mode,T,n,N,causal,expect", ATTN64_CASES)
def test_specialised_attention_kernels_on_the_cpu_simt_stand_in(hattn64, B, H, mode, T, n, N, causal, expect):
    d = 64
    torch.manual_seed(N + mode)
    qkv = torch.randn(B, N, 3 * H * d).to(BF16)
    dout = torch.randn(B * N, H * d).to(BF16)
    scale = d ** -0.5
    out = torch.full((B * N, H * d), float("nan"), dtype=BF16)
    lse = torch.full((B, H, N), float("nan"))
    kinds = ctypes.c_int(0)
    hattn64.h64_attn_fwd(P(qkv), P(out), P(lse), I(B), I(N), I(H), I(mode), I(T), I(n), I(int(causal)), ctypes.c_float(scale), ctypes.byref(kinds))
    assert kinds.value == expect
    ro, rl = torch.empty_like(out), torch.empty_like(lse)
    emu.attn_fwd(qkv, ro, rl, B, N, H, d, mode, T, n, int(causal), scale)
    assert torch.allclose(out.float(), ro.float(), atol=2e-2) and torch.allclose(lse, rl, atol=1e-4)
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty_like(lse)
    hattn64.h64_attn_bwd(P(qkv), P(ro), P(dout), P(rl), P(delta), P(dqkv), I(B), I(N), I(H), I(mode), I(T), I(n), I(int(causal)), ctypes.c_float(scale))
    rd, rdel = torch.empty_like(qkv), torch.empty_like(lse)
    emu.attn_bwd(qkv, ro, dout, rl, rdel, rd, B, N, H, d, mode, T, n, int(causal), scale)
    assert torch.isfinite(dqkv.float()).all()
    assert torch.allclose(dqkv.float(), rd.float(), atol=3e-2, rtol=3e-2), (dqkv.float() - rd.float()).abs().max()


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.fixture(scope="module")
def hln(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostln") / "libhostln.so")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-DTVTS_HOST_SHIM", "-I", HERE,
                    os.path.join(HERE, "harness_ln.cpp"), "-o", out], check=True)
    return ctypes.CDLL(out)


@pytest.mark.parametrize("M,D,grid", [(37, 640, 2), (21, 768, 3), (9, 128, 1), (12, 1280, 5)])
def test_layernorm_kernels_on_the_cpu_simt_stand_in(hln, M, D, grid):
    """layernorm.cu on the stand-in: width 640 (NV = 5, added for the H/14 test model) and the backward with the fused bias-gradient
    column sums (`dxsum`), neither of which has been through a GPU test yet; 768 / 128 / 1280 are the GPU-verified widths."""
    g_ = torch.Generator().manual_seed(M + D)
    x = torch.randn(M, D, generator=g_) * 2 + 0.5
    gam, bet = 1 + 0.1 * torch.randn(D, generator=g_), 0.1 * torch.randn(D, generator=g_)
    for y_bf16 in (1, 0):
        y = torch.empty(M, D, dtype=BF16 if y_bf16 else torch.float32)
        mean, rstd = torch.empty(M), torch.empty(M)
        assert hln.h_ln_fwd(P(x), P(gam), P(bet), P(y), ctypes.c_int(y_bf16), P(mean), P(rstd), I(M), I(D), ctypes.c_float(1e-5)) == 0
        ry = torch.empty_like(y)
        rm, rr = torch.empty(M), torch.empty(M)
        emu.layernorm_fwd(x, gam, bet, ry, y_bf16, rm, rr, M, D, 1e-5)
        assert torch.allclose(y.float(), ry.float(), atol=2e-2 if y_bf16 else 2e-5)
        assert torch.allclose(mean, rm, atol=1e-5) and torch.allclose(rstd, rr, atol=1e-5, rtol=1e-4)
    for dy_bf16 in (1, 0):
        dy = torch.randn(M, D, generator=g_).to(BF16 if dy_bf16 else torch.float32)
        r1, r2 = torch.randn(M, D, generator=g_), torch.randn(M, D, generator=g_)
        res = []
        for host in (True, False):
            dx, dxb = torch.empty(M, D), torch.empty(M, D, dtype=BF16)
            dg, db, dxs = torch.zeros(D), torch.zeros(D), torch.ones(D)
            if host:
                assert hln.h_ln_bwd(P(dy), ctypes.c_int(dy_bf16), P(x), P(rm), P(rr), P(gam), P(r1), P(r2), P(dx), P(dxb), P(dg), P(db), P(dxs),
                                    I(M), I(D), ctypes.c_int(grid)) == 0
            else:
                emu.layernorm_bwd_colsum(dy, dy_bf16, x, rm, rr, gam, r1, r2, dx, dxb, dg, db, dxs, M, D)
            res.append((dx, dxb, dg, db, dxs))
        assert torch.allclose(res[0][0], res[1][0], atol=2e-4)
        assert torch.allclose(res[0][1].float(), res[1][1].float(), atol=5e-2)
        for i in (2, 3, 4):
            assert torch.allclose(res[0][i], res[1][i], atol=1e-3 * M ** 0.5), i
        assert torch.allclose(res[0][4], 1.0 + res[0][0].sum(0), atol=1e-3)        # dxsum accumulates the column sums of dx
