"""Per-kernel parity on the B200: every C-ABI op of libtvts_b200.so against its torch restatement (tests/emu.py) on the
same seeded inputs.  Integer / index work must be bit-exact; fp32 kernels to ~1e-5; bf16-output kernels to bf16 rounding."""
import pytest
import torch

import emu
from tvts_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16, F32 = L.OPERAND_DTYPE, torch.float32


def rnd(*shape, dtype=F32, scale=1.0, seed=None):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


def close(a, b, atol, rtol=0.0, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), f"{what}: max abs err {err}"


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # M, N, K, a_mn, b_mn, mode, out_bf16
    (128, 256, 64, 0, 0, "plain", False),
    (297, 384, 128, 0, 0, "plain", False),           # ragged M (TMA OOB rows), N not a multiple of the tile
    (300, 128, 512, 0, 0, "plain", True),            # BN=128 path
    (64, 4, 128, 0, 0, "plain", False),              # tiny N
    (1000, 768, 768, 0, 0, "bias_act_res", False),     # fp32 out: bias + QuickGELU + residual
    (1000, 768, 768, 0, 0, "bias_act_pre", True),      # bf16 out + bf16 pre-activation copy (the c_fc GEMM)
    (598, 384, 128, 0, 0, "bias", True),               # pair tiles whose second CTA is entirely out of range
    (598, 128, 512, 0, 0, "bias_act_res", False),
    (1000, 2304, 768, 0, 0, "bias", True),
    (515, 512, 2048, 0, 1, "dact", True),            # dgrad form: B read MN-major
    (768, 2304, 8192, 0, 0, "splitk", False),
    (256, 512, 256, 1, 1, "plain", False),           # wgrad form: both operands MN-major
    (384, 128, 2376, 1, 1, "splitk", False),
    (768, 3072, 6288, 1, 1, "splitk", False),
    (0, 0, 0, 0, 0, "empty", False),
]


GEMM_CASES += [
    (2000, 3072, 768, 0, 0, "bias_act_pre", True),     # the c_fc shape class: two outputs, several tiles per worker
    (2000, 3072, 768, 0, 1, "dact", True),             # the c_proj dgrad shape class: act' operand prefetched through the ring
]


@pytest.fixture
def epi_direct(request):
    """tvts_gemm_set_epilogue_direct: 0 = shared-memory boxes + TMA stores (default), 1 = results stored straight from registers"""
    L.lib().tvts_gemm_set_epilogue_direct(int(request.param))
    yield request.param
    L.lib().tvts_gemm_set_epilogue_direct(0)


@pytest.mark.parametrize("epi_direct", [0, 1], indirect=True)
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,mode,out_bf16", GEMM_CASES)
def test_gemm(M, N, K, a_mn, b_mn, mode, out_bf16, epi_direct):
    torch.manual_seed(M + N + K)
    if mode == "empty":
        with pytest.raises(RuntimeError):
            L.gemm(torch.empty(8, 8, device=DEV, dtype=BF16), torch.empty(8, 8, device=DEV, dtype=BF16),
                   torch.empty(8, 8, device=DEV), M=0, N=8, K=8, lda=8, ldb=8)
        return
    A, B = rnd(M, K, scale=0.5).to(BF16), rnd(N, K, scale=0.5).to(BF16)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    lda, ldb = (M if a_mn else K), (N if b_mn else K)
    kw = {}
    out = torch.zeros(M, N, device=DEV, dtype=BF16 if out_bf16 else F32)
    if mode in ("bias", "bias_act_res", "bias_act_pre"):
        kw["bias"] = rnd(N)
    if mode == "bias_act_res":
        kw.update(residual=rnd(M, N), act="quick_gelu")
    if mode == "bias_act_pre":
        kw.update(act="gelu", out_pre=torch.empty(M, N, device=DEV, dtype=BF16))
    if mode == "dact":
        kw.update(aux=rnd(M, N).to(BF16), dact="gelu")
    if mode == "splitk":
        out = rnd(M, N)
        kw.update(accumulate=True)
    ref = out.clone()
    kw_ref = dict(kw)
    if "out_pre" in kw:
        kw_ref["out_pre"] = torch.empty_like(kw["out_pre"])
    emu.gemm(a, b, ref, M=M, N=N, K=K, lda=lda, ldb=ldb, a_mn=a_mn, b_mn=b_mn, **kw_ref)
    L.gemm(a, b, out, M=M, N=N, K=K, lda=lda, ldb=ldb, a_mn=a_mn, b_mn=b_mn, **kw)
    torch.cuda.synchronize()
    scale = ref.float().abs().max().item()
    close(out, ref, atol=(1e-2 if out_bf16 else 2e-5) * max(scale, 1.0), what="gemm out")
    if "out_pre" in kw:
        close(kw["out_pre"], kw_ref["out_pre"], atol=1e-2 * max(scale, 1.0), what="gemm out_pre")


@pytest.mark.parametrize("N,ldo,out_bf16,res", [(72, 96, True, False), (88, 96, True, False), (44, 48, False, True), (36, 64, False, False),
                                                (200, 208, True, False)])
@pytest.mark.parametrize("epi_direct", [1, 0], indirect=True)
def test_gemm_direct_epilogue_edges(N, ldo, out_bf16, res, epi_direct):
    """Register-store epilogue (and, for comparison, the TMA-store one) at the right edge of the matrix: the last 32-column pass holds 8 / 24 valid bf16 or 12 / 4 valid fp32 columns
    (16-byte and 32-byte pieces), rows are ldo apart; nothing outside [M, N] may be written."""
    torch.manual_seed(N)
    M, K = 300, 128
    A, B = rnd(M, K, scale=0.5).to(BF16), rnd(N, K, scale=0.5).to(BF16)
    kw = dict(bias=rnd(N))
    if res:
        kw.update(residual=rnd(M, N), act="quick_gelu")
    out = torch.full((M + 3, ldo), 7.0, device=DEV, dtype=BF16 if out_bf16 else F32)
    ref = out.clone()
    emu.gemm(A, B, ref, M=M, N=N, K=K, lda=K, ldb=K, ldo=ldo, **kw)
    L.gemm(A, B, out, M=M, N=N, K=K, lda=K, ldb=K, ldo=ldo, **kw)
    torch.cuda.synchronize()
    close(out[:M, :N], ref[:M, :N], atol=(1e-2 if out_bf16 else 2e-5) * max(ref[:M, :N].float().abs().max().item(), 1.0), what="gemm out")
    assert torch.equal(out[:M, N:], ref[:M, N:]) and torch.equal(out[M:], ref[M:]), "written outside the matrix"


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("M,D,bf16_out", [(1000, 768, True), (77, 512, True), (33, 128, False), (5, 1280, True)])
def test_layernorm(M, D, bf16_out):
    torch.manual_seed(M)
    x, g, b = rnd(M, D, scale=2.0) + 0.5, 1 + 0.1 * rnd(D), 0.1 * rnd(D)
    outs = []
    for fn in (L.call, lambda n, *a: emu.OPS[n](*a)):
        y = torch.empty(M, D, device=DEV, dtype=BF16 if bf16_out else F32)
        mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
        fn("layernorm_fwd", x, g, b, y, int(bf16_out), mean, rstd, M, D, 1e-5)
        outs.append((y, mean, rstd))
    close(outs[0][0], outs[1][0], atol=2e-2 if bf16_out else 2e-5, what="ln y")
    close(outs[0][1], outs[1][1], atol=1e-5, what="ln mean")
    close(outs[0][2], outs[1][2], atol=1e-5, rtol=1e-4, what="ln rstd")
    mean, rstd = outs[1][1], outs[1][2]
    for dy_bf16 in (False, True):
        dy = rnd(M, D).to(BF16 if dy_bf16 else F32)
        r1, r2 = rnd(M, D), rnd(M, D)
        res = []
        for fn in (L.call, lambda n, *a: emu.OPS[n](*a)):
            dx, dxb = torch.empty(M, D, device=DEV), torch.empty(M, D, device=DEV, dtype=BF16)
            dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
            fn("layernorm_bwd", dy, int(dy_bf16), x, mean, rstd, g, r1, r2, dx, dxb, dg, db, M, D)
            res.append((dx, dxb, dg, db))
        close(res[0][0], res[1][0], atol=2e-4, what="ln dx")
        close(res[0][1], res[1][1], atol=5e-2, what="ln dx bf16")
        close(res[0][2], res[1][2], atol=1e-3 * M ** 0.5, what="ln dgamma")
        close(res[0][3], res[1][3], atol=1e-3 * M ** 0.5, what="ln dbeta")


# ------------------------------------------------------------------------------------------------ attention
ATTN_CASES = [
    # B, H, mode, T, n, N, causal
    (2, 2, 0, 0, 0, 77, True),      # text tower
    (2, 2, 0, 0, 0, 200, False),    # sort head, several streamed tiles
    (1, 3, 0, 0, 0, 64, False),
    (2, 2, 1, 2, 49, 99, False),    # space, B/32 c1
    (2, 2, 2, 2, 49, 99, False),    # time
    (1, 2, 1, 3, 98, 295, False),   # space, B/16 masked (2 stationary chunks, 2 streamed tiles)
    (1, 2, 2, 8, 5, 41, False),     # time, T=8
    (2, 3, 2, 12, 7, 85, False),    # time, T=12 (the shipped 3x4-frame configs)
    (1, 2, 2, 16, 6, 97, False),    # time, T=16: 17 keys -> third key tile / second key m-tile
    (1, 2, 2, 20, 3, 61, False),    # time, T>16: streamed fallback with strided groups
    (1, 2, 1, 2, 196, 393, False),  # space, unmasked B/16 frame: 4 stationary chunks, 4 streamed tiles
    (2, 2, 1, 8, 98, 785, False),   # space, c3 shape
    (2, 2, 2, 8, 98, 785, False),   # time, c3 shape
    (1, 2, 0, 0, 0, 789, False),    # sort head at c3: 13 chunks x 13 tiles
    (3, 2, 0, 0, 0, 130, True),     # causal across several tiles (dK/dV start tile > 0)
]


@pytest.fixture
def attn_tc(request):
    """tvts_attn_set_tc: 3 = tcgen05 one-tile kernels (csrc/attention_tc.cu) for space / short sequences AND time attention (default),
    1 = tcgen05 for space / short sequences only, 0 = the mma.sync kernels for every shape"""
    L.lib().tvts_attn_set_tc(int(request.param))
    yield request.param
    L.lib().tvts_attn_set_tc(3)


ATTN_CASES_TC = ATTN_CASES + [
    (3, 12, 1, 8, 98, 785, False),  # space at the c3 head count: 288 tcgen05 tiles + the CLS merge over 8 frames
    (2, 8, 0, 0, 0, 77, True),      # text tower heads
    (2, 2, 0, 0, 0, 128, True),     # a full 128-row tile (no padding rows), causal
    (2, 2, 0, 0, 0, 17, False),     # a nearly empty tile
    (1, 2, 1, 1, 127, 128, False),  # the largest space group (127 patches + CLS)
    (2, 2, 1, 3, 5, 16, False),     # tiny space groups
    (1, 2, 2, 8, 23, 185, False),   # time: 2 tiles of 12 positions, the last one ragged (11 positions + 1 zero-filled)
    (1, 2, 2, 3, 51, 154, False),   # time: 2 tiles of 26 / 25 positions x 3 frames
    (2, 12, 2, 8, 98, 785, False),  # time at the c3 head count: 7 tiles of 14 positions x 8 frames per (b, h)
    (1, 2, 2, 1, 30, 31, False),    # time with a single frame (every group = one patch + CLS)
]


@pytest.mark.parametrize("attn_tc", [3, 1, 0], indirect=True)
@pytest.mark.parametrize("B,H,mode,T,n,N,causal", ATTN_CASES_TC)
def test_attention(B, H, mode, T, n, N, causal, attn_tc):
    torch.manual_seed(N + mode)
    d = 64
    qkv = rnd(B, N, 3 * H * d, scale=1.0).to(BF16)
    dout = rnd(B * N, H * d).to(BF16)
    scale = d ** -0.5
    res = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        out = torch.empty(B * N, H * d, device=DEV, dtype=BF16)
        lse = torch.empty(B, H, N, device=DEV)
        fn("attn_fwd", qkv, out, lse, B, N, H, d, mode, T, n, int(causal), scale)
        res.append((out, lse))
    close(res[0][0], res[1][0], atol=2e-2, what="attn out")
    close(res[0][1], res[1][1], atol=1e-3, what="attn lse")
    out, lse = res[1]
    grads = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        dqkv = torch.full_like(qkv, float("nan"))
        delta = torch.empty_like(lse)
        fn("attn_bwd", qkv, out, dout, lse, delta, dqkv, B, N, H, d, mode, T, n, int(causal), scale)
        grads.append(dqkv)
    assert torch.isfinite(grads[0].float()).all(), "attn_bwd left elements unwritten"
    close(grads[0], grads[1], atol=3e-2, rtol=3e-2, what="attn dqkv")
    # the variant that also accumulates the qkv bias gradient (column sums of dqkv): same dqkv bit for bit, bias = sums of the stored values
    dqkv2 = torch.full_like(qkv, float("nan"))
    dbias = torch.full((3 * H * d,), 0.5, device=DEV)
    L.call("attn_bwd_bias", qkv, out, dout, lse, torch.empty_like(lse), dqkv2, dbias, B, N, H, d, mode, T, n, int(causal), scale)
    assert torch.equal(dqkv2, grads[0])
    want = 0.5 + grads[0].view(B * N, 3 * H * d).float().sum(0)
    close(dbias, want, atol=2e-3 * max(1.0, want.abs().max().item()), what="attn qkv bias gradient")


@pytest.mark.parametrize("B,H,N,q0,qn", [(2, 2, 103, 99, 4), (2, 8, 789, 785, 4), (1, 2, 300, 100, 70)])
def test_attention_query_window(B, H, N, q0, qn):
    """Only rows [q0, q0+qn) are queries (last sort-head block): fwd rows, dq of the window, dk/dv of every token."""
    torch.manual_seed(N)
    d = 64
    qkv = rnd(B, N, 3 * H * d, scale=1.0).to(BF16)
    scale = d ** -0.5
    res = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        out = torch.zeros(B * N, H * d, device=DEV, dtype=BF16)
        lse = torch.zeros(B, H, N, device=DEV)
        fn("attn_window_fwd", qkv, out, lse, B, N, H, d, q0, qn, scale)
        res.append((out, lse))
    close(res[0][0], res[1][0], atol=2e-2, what="window out")
    close(res[0][1], res[1][1], atol=1e-3, what="window lse")
    out, lse = res[1]
    dout = torch.zeros(B, N, H * d, device=DEV, dtype=BF16)
    dout[:, q0:q0 + qn] = rnd(B, qn, H * d).to(BF16)
    grads = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        dqkv = torch.zeros_like(qkv)
        delta = torch.empty_like(lse)
        fn("attn_window_bwd", qkv, out, dout.view(B * N, H * d), lse, delta, dqkv, B, N, H, d, q0, qn, scale)
        grads.append(dqkv)
    close(grads[0], grads[1], atol=3e-2, rtol=3e-2, what="window dqkv")


# ------------------------------------------------------------------------------------------------ glue kernels
def both(name, make):
    """run op `name` with the real kernel and the restatement on identical fresh argument lists; return the two lists"""
    torch.manual_seed(0)
    a1 = make()
    torch.manual_seed(0)
    a2 = make()
    L.call(name, *a1)
    emu.OPS[name](*a2)
    torch.cuda.synchronize()
    return a1, a2


def test_cast_and_colsum():
    a1, a2 = both("cast_bf16", lambda: [rnd(1001), torch.empty(1001, device=DEV, dtype=BF16), 1001])
    assert torch.equal(a1[1], a2[1])
    a1, a2 = both("colsum_bf16", lambda: [rnd(3000, 2304).to(BF16), torch.zeros(2304, device=DEV), 3000, 2304, 2304])
    close(a1[1], a2[1], atol=2e-2, what="colsum")


@pytest.mark.parametrize("B,T,R,p,n", [(2, 2, 224, 32, 49), (2, 3, 224, 16, 98), (1, 1, 64, 16, 7)])
def test_patch_gather_and_assemble(B, T, R, p, n):
    P, D = (R // p) ** 2, 128
    keep = torch.stack([torch.randperm(P, device=DEV)[:n] for _ in range(B)]).contiguous()
    a1, a2 = both("patch_gather", lambda: [rnd(B, T, 3, R, R), keep, torch.empty(B * T * n, 3 * p * p, device=DEV, dtype=BF16), B, T, R, p, n])
    assert torch.equal(a1[2], a2[2]), "patch gather must be bit-exact (pure indexing + rounding)"
    mk = lambda: [rnd(B * T * n, D), rnd(D), rnd(P + 1, D), rnd(T + 2, D), keep, torch.empty(B * (1 + T * n), D, device=DEV), B, T, n, D]
    a1, a2 = both("video_assemble", mk)
    close(a1[5], a2[5], atol=1e-6, what="video_assemble")
    mk = lambda: [rnd(B * (1 + T * n), D), keep, torch.zeros(D, device=DEV), torch.zeros(P + 1, D, device=DEV),
                  torch.zeros(T + 2, D, device=DEV), torch.empty(B * T * n, D, device=DEV, dtype=BF16), B, T, n, D]
    a1, a2 = both("video_assemble_bwd", mk)
    for i, w in ((2, "dcls"), (3, "dpos"), (4, "dtem")):
        close(a1[i], a2[i], atol=1e-4, what=w)
    assert torch.equal(a1[5], a2[5])


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
def test_text_embed_argmax_gather(dtype):
    rows, Lc, W, V = 12, 77, 128, 512
    from tvts_b200 import config as C
    from tvts_b200.synthetic import make_tokens
    tok = make_tokens(C.TINY_B, rows, seed=3, dtype=dtype).to(DEV)
    is64 = int(dtype == torch.int64)
    a1, a2 = both("text_embed", lambda: [tok, is64, rnd(V, W), rnd(Lc, W), torch.empty(rows * Lc, W, device=DEV), rows, Lc, W])
    assert torch.equal(a1[4], a2[4])
    a1, a2 = both("argmax_rows", lambda: [tok, is64, torch.empty(rows, dtype=torch.int64, device=DEV), rows, Lc])
    assert torch.equal(a1[2], a2[2]), "EOT index must be exact"
    idx = a2[2]
    a1, a2 = both("text_embed_bwd", lambda: [rnd(rows * Lc, W), tok, is64, torch.zeros(V, W, device=DEV), torch.zeros(Lc, W, device=DEV), rows, Lc, W])
    close(a1[3], a2[3], atol=1e-4, what="dtable")
    close(a1[4], a2[4], atol=1e-4, what="dpos")
    a1, a2 = both("gather_rows", lambda: [rnd(rows * Lc, W), idx, torch.empty(rows, W, device=DEV), rows, W])
    assert torch.equal(a1[2], a2[2])
    a1, a2 = both("scatter_rows", lambda: [rnd(rows, W), idx, torch.zeros(rows * Lc, W, device=DEV), rows, W, 0])
    assert torch.equal(a1[2], a2[2])


def test_group_mean_sort_concat_small_linear():
    nt, B, E, N = 4, 3, 128, 9
    a1, a2 = both("group_mean", lambda: [rnd(nt * B, E), torch.empty(B, E, device=DEV), nt, B, E])
    close(a1[1], a2[1], atol=1e-6)
    a1, a2 = both("group_mean_bwd", lambda: [rnd(B, E), torch.empty(nt * B, E, device=DEV), None, nt, B, E])
    close(a1[1], a2[1], atol=1e-6)
    a1, a2 = both("sort_concat", lambda: [rnd(B, N, E), rnd(nt * B, E), rnd(2, E), torch.empty(B * (N + nt), E, device=DEV), B, N, nt, E])
    assert torch.equal(a1[3], a2[3])
    a1, a2 = both("sort_concat_bwd", lambda: [rnd(B * (N + nt), E), torch.empty(B, N, E, device=DEV), torch.zeros(2, E, device=DEV), B, N, nt, E])
    assert torch.equal(a1[1], a2[1])
    close(a1[2], a2[2], atol=1e-4)
    R, K, O = B * nt, E, 4
    a1, a2 = both("small_linear_fwd", lambda: [rnd(R, K), rnd(O, K), rnd(O), torch.empty(R, O, device=DEV), R, K, O])
    close(a1[3], a2[3], atol=1e-4)
    a1, a2 = both("small_linear_bwd", lambda: [rnd(R, O), rnd(R, K), rnd(O, K), torch.empty(R, K, device=DEV), torch.zeros(O, K, device=DEV),
                                                torch.zeros(O, device=DEV), R, K, O])
    for i in (3, 4, 5):
        close(a1[i], a2[i], atol=1e-4)
    a1, a2 = both("add_rows", lambda: [rnd(B, 2 * E), rnd(B, 3 * E), B, E, 2 * E, 3 * E])
    close(a1[1], a2[1], atol=1e-6)


@pytest.mark.parametrize("Bg,E", [(8, 128), (256, 512), (37, 512)])
def test_losses(Bg, E):
    from tvts_b200 import engine as Eng
    torch.manual_seed(Bg)
    v = rnd(Bg, E).requires_grad_(True)
    t = (rnd(Bg, E) + 0.5 * v.detach()).requires_grad_(True)
    loss = Eng.norm_softmax_loss(Eng.sim_matrix(v, t), 0.05)
    loss.backward()
    v2, t2 = v.detach().clone().requires_grad_(True), t.detach().clone().requires_grad_(True)
    an = v2 / v2.norm(dim=1, keepdim=True).clamp_min(1e-8)
    bn = t2 / t2.norm(dim=1, keepdim=True).clamp_min(1e-8)
    z = an @ bn.t() / 0.05
    ref = -torch.diagonal(torch.log_softmax(z, 1)).mean() - torch.diagonal(torch.log_softmax(z.t(), 1)).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    close(v.grad, v2.grad, atol=1e-5 + 1e-3 * v2.grad.abs().max().item(), what="d video")
    close(t.grad, t2.grad, atol=1e-5 + 1e-3 * t2.grad.abs().max().item(), what="d text")
    # sort CE: labels are exact integers
    R, Cc = 4 * Bg, 4
    x = rnd(R // 4, 4, Cc).requires_grad_(True)
    y = torch.arange(4, device=DEV).repeat(R // 4, 1)
    l2 = Eng.sort_ce(x, y)
    l2.backward()
    x2 = x.detach().clone().requires_grad_(True)
    r2 = 2 * torch.nn.functional.cross_entropy(x2.reshape(-1, Cc), y.reshape(-1))
    r2.backward()
    assert abs(l2.item() - r2.item()) < 1e-5 * max(1.0, r2.item())
    close(x.grad, x2.grad, atol=1e-6)


@pytest.mark.parametrize("Bg,E,row0,nloc,with_ce", [(8, 128, 0, 8, True), (256, 512, 96, 32, True), (37, 512, 5, 9, True), (32, 512, 0, 32, True),
                                                      (64, 1024, 32, 32, False), (256, 512, 0, 256, True), (16, 128, 8, 8, True)])
def test_fused_contrastive_sortce(Bg, E, row0, nloc, with_ce):
    """csrc/loss_fused.cu (ONE launch, cluster of up to 8 CTAs) against the torch restatement of sim_matrix + NormSoftmaxLoss + 2 * CE:
    losses and the gradient of the LOCAL rows of both gathered embeddings; sort labels are exact integers."""
    torch.manual_seed(Bg + E)
    v = rnd(Bg, E)
    t = rnd(Bg, E) + 0.5 * v
    v[Bg // 2] *= 1e-12                          # a row below the eps clamp of sim_matrix
    R, C = 4 * nloc, 4
    x = rnd(R, C) if with_ce else None
    y = torch.arange(4, device=DEV).repeat(nloc) if with_ce else None
    outs = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        loss1, loss2 = torch.full((), float("nan"), device=DEV), torch.full((), float("nan"), device=DEV)
        dv, dt = torch.full((nloc, E), float("nan"), device=DEV), torch.full((nloc, E), float("nan"), device=DEV)
        dx = torch.full((R, C), float("nan"), device=DEV) if with_ce else None
        fn("contrastive_sortce_fused", v, t, Bg, E, row0, nloc, 0.05, 1e-8, x, y, R if with_ce else 0, C if with_ce else 0, 2.0,
           loss1, loss2 if with_ce else None, dv, dt, dx)
        outs.append((loss1, loss2, dv, dt, dx))
    a, b = outs
    assert abs(a[0].item() - b[0].item()) < 1e-4 * max(1.0, abs(b[0].item())), (a[0].item(), b[0].item())
    close(a[2], b[2], atol=1e-5 + 1e-3 * b[2].abs().max().item(), what="d video (local rows)")
    close(a[3], b[3], atol=1e-5 + 1e-3 * b[3].abs().max().item(), what="d text (local rows)")
    if with_ce:
        assert abs(a[1].item() - b[1].item()) < 1e-5 * max(1.0, b[1].item())
        close(a[4], b[4], atol=1e-6, what="d logits")


def test_fused_losses_autograd_matches_the_separate_kernels():
    """engine.fused_losses (what TrainStep uses) against engine.sim_matrix + norm_softmax_loss + sort_ce: same losses, same gradients."""
    from tvts_b200 import engine as Eng
    from tvts_b200.trainer import gather_for_fused_loss
    torch.manual_seed(3)
    B, E, nt = 6, 128, 4
    v0, t0, p0 = rnd(B, E), rnd(B, E), rnd(B, nt, nt)
    y = torch.arange(nt, device=DEV).repeat(B, 1)
    res = []
    for fused in (True, False):
        v, t, p = (z.clone().requires_grad_(True) for z in (v0, t0, p0))
        if fused:
            l1, l2 = Eng.fused_losses(v, t, p, y, 0.05, gather_for_fused_loss)
        else:
            l1, l2 = Eng.norm_softmax_loss(Eng.sim_matrix(v, t), 0.05), Eng.sort_ce(p, y, 2.0)
        ((l1 + l2) * 8.0).backward()
        res.append((l1.item(), l2.item(), v.grad, t.grad, p.grad))
    assert abs(res[0][0] - res[1][0]) < 1e-5 and abs(res[0][1] - res[1][1]) < 1e-5
    for i in (2, 3, 4):
        close(res[0][i], res[1][i], atol=1e-6 + 1e-4 * res[1][i].abs().max().item(), what=f"grad {i}")
