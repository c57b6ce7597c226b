"""GPU parity tests of the H/14 (c4), TVTS v1 (c5), uint8-input, text-trimming and downstream paths (first executed on a B200 in round 2,
GPU call 1: 87 of 87 cases green -- profiles/r2/call1_staged_suite_87_passed.txt):
  * the head-dim-generic streamed attention kernels (tvts_b200/csrc/attention_hd.cu): d = 80 (ViT-H/14) against the torch
    restatement, d = 64 against the restatement AND against the specialised kernels of attention.cu
  * the padded patch-embed path of 14x14 patches (patch_gather_ld, cast_bf16_pad, GEMMs with K = 592 / N = 588)
  * LayerNorm width 640, the TVTSv2_H_14 model (tiny_H640) against the executed-reference fixture and the oracle's gradients
  * the downstream (zero-shot) towers against their executed-reference fixture
  * the fused uint8 input stage (tvts_b200/csrc/input_stage.cu): bit-exact against the float path
  * text-context trimming (trainer.trim_text_context): same results from a token matrix cut to the batch's longest caption
  * TVTS v1: tubelet gather / per-tube assembly / ReLU kernels (tvts_b200/csrc/v1_glue.cu), key-padded attention, and the whole v1 model
    (DistilBERT text encoder included) against the executed-reference fixture tiny_v1_full and the oracle's gradients
Tolerances as in tests/test_kernels_gpu.py / tests/test_model_gpu.py."""
import os
import types

import numpy as np
import pytest
import torch

import emu
import tvts_oracle as O
from tvts_b200 import _lib as L
from tvts_b200 import config as C
from tvts_b200 import engine as E
from tvts_b200 import modules as M
from tvts_b200.synthetic import make_batch, make_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16, F32 = L.OPERAND_DTYPE, torch.float32
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rnd(*shape, scale=1.0):
    return torch.randn(*shape, device=DEV) * scale


def close(a, b, atol, rtol=0.0, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), f"{what}: max abs err {err}"


def to_cuda(data):
    return {k: (v.cuda() if k != "keep_ind" else v) for k, v in data.items()}


GENERIC_ATTN_CASES = [
    # B, H, mode, T, n, N, causal
    (2, 2, 0, 0, 0, 77, True),      # causal, two streamed tiles
    (2, 2, 0, 0, 0, 200, False),    # full, several tiles
    (1, 3, 0, 0, 0, 64, False),
    (2, 2, 1, 2, 49, 99, False),    # space
    (2, 2, 2, 2, 49, 99, False),    # time (streamed: strided groups)
    (1, 2, 1, 3, 76, 229, False),   # space, the H/14 kept-patch count (int(256 * 0.3) = 76): 2 stationary chunks
    (1, 2, 2, 3, 76, 229, False),   # time
    (1, 2, 2, 16, 6, 97, False),    # time, c4's 16 frames
    (1, 2, 1, 16, 76, 1217, False),  # space at c4's token count
    (3, 2, 0, 0, 0, 130, True),     # causal across tiles
    (1, 2, 1, 2, 100, 201, False),  # space, 101 rows per group (7 warps)
    (2, 2, 0, 0, 0, 40, False),     # short full-attention sequence (group-resident in mode 0)
    (1, 2, 2, 12, 76, 913, False),  # time attention at the shipped H/14 clip length: warp-per-slot kernels
    (1, 2, 2, 15, 6, 91, False),    # the largest T of the one-tile warp-per-slot kernels
    (1, 2, 2, 16, 76, 1217, False),  # time attention of c4 (16 frames): two row tiles per slot
    (1, 2, 2, 31, 2, 63, False),    # the largest T of the two-tile kernels
    (1, 2, 2, 32, 2, 65, False),    # beyond it: streamed (strided groups)
]


@pytest.fixture
def hd_group(request):
    """tvts_attn_hd_set_group: group-resident generic kernels allowed (default) or streamed kernels only"""
    L.lib().tvts_attn_hd_set_group(int(request.param))
    yield request.param
    L.lib().tvts_attn_hd_set_group(1)


@pytest.mark.parametrize("hd_group", [1, 0], indirect=True)
@pytest.mark.parametrize("d", [80, 64])
@pytest.mark.parametrize("B,H,mode,T,n,N,causal", GENERIC_ATTN_CASES)
def test_generic_attention(B, H, mode, T, n, N, causal, d, hd_group):
    torch.manual_seed(N + mode + d)
    qkv = rnd(B, N, 3 * H * d).to(BF16)
    dout = rnd(B * N, H * d).to(BF16)
    scale = d ** -0.5
    res = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        out = torch.empty(B * N, H * d, device=DEV, dtype=BF16)
        lse = torch.empty(B, H, N, device=DEV)
        fn("attn_generic_fwd", qkv, out, lse, B, N, H, d, mode, T, n, int(causal), scale)
        res.append((out, lse))
    close(res[0][0], res[1][0], atol=2e-2, what="attn out")
    close(res[0][1], res[1][1], atol=1e-3, what="attn lse")
    out, lse = res[1]
    grads = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        dqkv = torch.full_like(qkv, float("nan"))
        delta = torch.empty_like(lse)
        fn("attn_generic_bwd", qkv, out, dout, lse, delta, dqkv, B, N, H, d, mode, T, n, int(causal), scale)
        grads.append(dqkv)
    assert torch.isfinite(grads[0].float()).all(), "attn_generic_bwd left elements unwritten"
    close(grads[0], grads[1], atol=3e-2, rtol=3e-2, what="attn dqkv")
    if d == 80:     # the public entry points forward d != 64 to the generic kernels
        out2, lse2 = torch.empty_like(out), torch.empty_like(lse)
        L.call("attn_fwd", qkv, out2, lse2, B, N, H, d, mode, T, n, int(causal), scale)
        assert torch.equal(out2, res[0][0]) and torch.equal(lse2, res[0][1])
    else:           # same math as the specialised head-dim-64 kernels (different tiling of the groups: bf16 rounding only)
        out2, lse2 = torch.empty_like(out), torch.empty_like(lse)
        L.call("attn_fwd", qkv, out2, lse2, B, N, H, d, mode, T, n, int(causal), scale)
        close(out2, res[0][0], atol=1e-2, what="generic<64> vs specialised out")
        close(lse2, res[0][1], atol=1e-4, what="generic<64> vs specialised lse")


@pytest.mark.parametrize("B,T,R,p,n", [(2, 3, 224, 14, 76), (1, 2, 56, 14, 5)])
def test_padded_patch_gather_and_weight_cast(B, T, R, p, n):
    P = (R // p) ** 2
    K = 3 * p * p
    Kp = (K + 7) // 8 * 8
    torch.manual_seed(p + n)
    keep = torch.stack([torch.randperm(P, device=DEV)[:n] for _ in range(B)]).contiguous()
    video = rnd(B, T, 3, R, R)
    got = torch.full((B * T * n, Kp), 7.0, device=DEV, dtype=BF16)
    ref = torch.full((B * T * n, Kp), 7.0, device=DEV, dtype=BF16)
    L.call("patch_gather_ld", video, keep, got, B, T, R, p, n, Kp)
    emu.patch_gather_ld(video, keep, ref, B, T, R, p, n, Kp)
    assert torch.equal(got, ref), "padded patch gather must be bit-exact (indexing + rounding, zero tail)"
    w = rnd(640, 3, p, p)
    got = torch.full((640, Kp), 7.0, device=DEV, dtype=BF16)
    ref = torch.full((640, Kp), 7.0, device=DEV, dtype=BF16)
    L.call("cast_bf16_pad", w, got, 640, K, Kp)
    emu.cast_bf16_pad(w, ref, 640, K, Kp)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("B,T,R,p,n", [(2, 3, 224, 16, 98), (2, 2, 224, 32, 49), (1, 1, 64, 16, 7)])
def test_fused_uint8_input_stage(B, T, R, p, n):
    """uint8 crops -> x/255 -> (x-mean)/std -> bf16 im2col of the kept patches, in one kernel: bit-identical to the float path."""
    import ctypes
    P = (R // p) ** 2
    torch.manual_seed(R + p)
    keep = torch.stack([torch.randperm(P, device=DEV)[:n] for _ in range(B)]).contiguous()
    u8 = torch.randint(0, 256, (B, T, 3, R, R), device=DEV, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    got = torch.empty(B * T * n, 3 * p * p, device=DEV, dtype=BF16)
    ref = torch.empty_like(got)
    L.call("patch_gather_u8", u8, keep, got, B, T, R, p, n, (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std))
    emu.patch_gather_u8(u8, keep, ref, B, T, R, p, n, mean, std)
    assert torch.equal(got, ref), (got.float() - ref.float()).abs().max()
    via_float = torch.empty_like(got)
    m = torch.tensor(mean, device=DEV)[None, None, :, None, None]
    s_ = torch.tensor(std, device=DEV)[None, None, :, None, None]
    L.call("patch_gather", u8.float().div(255).sub(m).div(s_).contiguous(), keep, via_float, B, T, R, p, n)
    assert torch.equal(got, via_float)


def test_patch_embed_gemms_with_padded_rows():
    """forward: tok[M, D] = cols[M, 592] . w[D, 592]^T ; wgrad: dW[D, 588] += dtok[M, D]^T . cols[M, :588] (operand pitch 592)."""
    torch.manual_seed(5)
    Mrows, D, K, Kp = 456, 640, 588, 592
    cols = torch.zeros(Mrows, Kp, device=DEV, dtype=BF16)
    cols[:, :K] = rnd(Mrows, K, scale=0.5).to(BF16)
    w = torch.zeros(D, Kp, device=DEV, dtype=BF16)
    w[:, :K] = rnd(D, K, scale=0.5).to(BF16)
    out, ref = torch.zeros(Mrows, D, device=DEV), torch.zeros(Mrows, D, device=DEV)
    L.gemm(cols, w, out, M=Mrows, N=D, K=Kp, lda=Kp, ldb=Kp)
    emu.gemm(cols, w, ref, M=Mrows, N=D, K=Kp, lda=Kp, ldb=Kp)
    close(out, ref, atol=2e-5 * max(ref.abs().max().item(), 1.0), what="padded patch-embed fwd")
    dtok = rnd(Mrows, D, scale=0.5).to(BF16)
    dw0 = rnd(D, K)
    dw, dref = dw0.clone(), dw0.clone()
    L.gemm(dtok, cols, dw, M=D, N=K, K=Mrows, lda=D, ldb=Kp, a_mn=True, b_mn=True, accumulate=True)
    emu.gemm(dtok, cols, dref, M=D, N=K, K=Mrows, lda=D, ldb=Kp, a_mn=True, b_mn=True, accumulate=True)
    close(dw, dref, atol=2e-5 * max(dref.abs().max().item(), 1.0), what="padded patch-embed wgrad")


def test_layernorm_width_640():
    Mrows, D = 333, 640
    torch.manual_seed(1)
    x, g, b = rnd(Mrows, D, scale=2.0) + 0.5, 1 + 0.1 * rnd(D), 0.1 * rnd(D)
    outs = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        y = torch.empty(Mrows, D, device=DEV, dtype=BF16)
        mean, rstd = torch.empty(Mrows, device=DEV), torch.empty(Mrows, device=DEV)
        fn("layernorm_fwd", x, g, b, y, 1, mean, rstd, Mrows, D, 1e-5)
        outs.append((y, mean, rstd))
    close(outs[0][0], outs[1][0], atol=2e-2, what="ln y")
    close(outs[0][2], outs[1][2], atol=1e-5, rtol=1e-4, what="ln rstd")
    mean, rstd = outs[1][1], outs[1][2]
    dy, r1 = rnd(Mrows, D).to(BF16), rnd(Mrows, D)
    res = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        dx, dxb = torch.empty(Mrows, D, device=DEV), torch.empty(Mrows, D, device=DEV, dtype=BF16)
        dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
        fn("layernorm_bwd", dy, 1, x, mean, rstd, g, r1, None, dx, dxb, dg, db, Mrows, D)
        res.append((dx, dg, db))
    close(res[0][0], res[1][0], atol=2e-4, what="ln dx")
    close(res[0][1], res[1][1], atol=1e-3 * Mrows ** 0.5, what="ln dgamma")
    close(res[0][2], res[1][2], atol=1e-3 * Mrows ** 0.5, what="ln dbeta")


def test_h14_model_against_reference_golden_and_oracle():
    """TVTSv2_H_14 at the smallest width the kernels take (tiny_H640: 8 heads x 80, 14x14 patches, mask 0.7, exact GELU, ln_post on
    CLS only, sort head over the patch tokens); fixture from the executed reference, gradients from the oracle."""
    cfg = C.TINY_H640
    g = np.load(os.path.join(GOLD, "tiny_H640.npz"))
    m = M.TVTSv2_H_14(types.SimpleNamespace(local_rank=0), arch=cfg)
    sd = make_state_dict(cfg, seed=1234)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    dd = to_cuda(data)
    te, ve, pred = m(dd)
    loss1 = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te))
    loss2 = E.sort_ce(pred, dd["label"])
    (loss1 + loss2).backward()
    torch.cuda.synchronize()
    assert abs(loss1.item() - float(g["loss1"])) < 5e-2 and abs(loss2.item() - float(g["loss2"])) < 5e-2
    np.testing.assert_allclose(te.detach().cpu().numpy(), g["text_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(ve.detach().cpu().numpy(), g["video_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["pred_order"], atol=5e-2, rtol=5e-2)
    _, _, _, ograds = O.step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg)
    got = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ograds)
    for k, gr in ograds.items():
        rel = (got[k].double() - gr.double()).norm().item() / (gr.double().norm().item() + 1e-8)
        assert rel < 0.08, (k, rel)


def test_downstream_model_against_reference_golden():
    """v2/downstream towers (mask_ratio 0, no sort head, forward only) on the CUDA kernels vs the executed-reference fixture."""
    g = np.load(os.path.join(GOLD, "tiny_ds.npz"))
    cfg = C.TINY_B
    sd = {k: v for k, v in make_state_dict(cfg, seed=1234).items() if not k.startswith("pred_model.")}
    data = make_batch(cfg, int(g["batch"]), int(g["frames"]), n_trans=int(g["n_trans"]), seed=int(g["seed"]))
    for cls, tag in ((M.TVTSv2_B_32_downstream, ""), (M.TVTSv2_B_32_downstream_mc, "_mc")):
        m = cls(arch=cfg)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            te, ve = m(to_cuda(data), return_embeds=True)
            np.testing.assert_allclose(te.cpu().numpy(), g["text_emb" + tag], atol=3e-2, rtol=3e-2)
            np.testing.assert_allclose(ve.cpu().numpy(), g["video_emb" + tag], atol=3e-2, rtol=3e-2)
            if tag == "":
                np.testing.assert_allclose(m(to_cuda(data), return_embeds=False).cpu().numpy(), g["sims"], atol=3e-2)


def test_text_context_trimming_on_the_gpu():
    """trainer.trim_text_context: the text tower on a token matrix cut to the longest caption of the batch (causal group-resident
    kernels at sequence lengths < 77) gives the same embeddings / losses / gradients as the full 77-column matrix."""
    from tvts_b200.trainer import trim_text_context
    cfg = C.TINY_B
    data = make_batch(cfg, 3, 2, n_trans=4, seed=21)
    trimmed = trim_text_context(data["text"])
    assert trimmed.shape[1] < cfg.context
    outs = []
    for text in (data["text"], trimmed):
        m = M.TVTSv2Base(types.SimpleNamespace(local_rank=0), arch=cfg)
        m.load_state_dict(make_state_dict(cfg, seed=1234), strict=True)
        m = m.cuda()
        dd = to_cuda(dict(data, text=text))
        te, ve, pred = m(dd)
        loss = M.NormSoftmaxLoss(cfg.temperature)(M.sim_matrix(ve, te)) + E.sort_ce(pred, dd["label"])
        loss.backward()
        torch.cuda.synchronize()
        outs.append((te.detach(), loss.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    close(outs[0][0], outs[1][0], atol=2e-3, what="text embeddings")          # same math per row; different kernel tiling only
    assert abs(outs[0][1].item() - outs[1][1].item()) < 2e-3
    for k, g in outs[0][2].items():
        rel = (g.double() - outs[1][2][k].double()).norm().item() / (g.double().norm().item() + 1e-8)
        assert rel < 2e-2, (k, rel)


# ------------------------------------------------------------------------------------------------ TVTS v1 (configs[4])
def both(name, make):
    torch.manual_seed(0)
    a1 = make()
    torch.manual_seed(0)
    a2 = make()
    L.call(name, *a1)
    emu.OPS[name](*a2)
    torch.cuda.synchronize()
    return a1, a2


@pytest.mark.parametrize("B,T,R,p,n", [(2, 8, 64, 16, 8), (1, 16, 224, 16, 49), (2, 4, 224, 32, 20)])
def test_v1_tubelet_gather_and_assemble(B, T, R, p, n):
    P, D, nt = (R // p) ** 2, 128, T // 2
    keep = torch.stack([torch.stack([torch.randperm(P, device=DEV)[:n] for _ in range(nt)]) for _ in range(B)]).contiguous()
    a1, a2 = both("tubelet_gather", lambda: [rnd(B, T, 3, R, R), keep, torch.empty(B * nt * n, 6 * p * p, device=DEV, dtype=BF16), B, T, R, p, n])
    assert torch.equal(a1[2], a2[2]), "tubelet gather must be bit-exact (pure indexing + rounding)"
    mk = lambda: [rnd(B * nt * n, D), rnd(1, 1, D), rnd(1, P + 1, D), rnd(1, nt + 1, D), keep, torch.empty(B * (1 + nt * n), D, device=DEV), B, nt, n, D]
    a1, a2 = both("video_assemble_tube", mk)
    close(a1[5], a2[5], atol=1e-6, what="video_assemble_tube")
    mk = lambda: [rnd(B * (1 + nt * n), D), keep, torch.zeros(1, 1, D, device=DEV), torch.zeros(1, P + 1, D, device=DEV),
                  torch.zeros(1, nt + 1, D, device=DEV), torch.empty(B * nt * n, D, device=DEV, dtype=BF16), B, nt, n, D]
    a1, a2 = both("video_assemble_tube_bwd", mk)
    for i, w in ((2, "dcls"), (3, "dpos"), (4, "dtem")):
        close(a1[i], a2[i], atol=1e-4, what=w)
    assert torch.equal(a1[5], a2[5])


def test_v1_relu_kernels():
    a1, a2 = both("relu_bf16", lambda: [rnd(8, 768), torch.empty(8, 768, device=DEV, dtype=BF16), 8 * 768])
    assert torch.equal(a1[1], a2[1])
    a1, a2 = both("relu_bwd", lambda: [rnd(8, 768), rnd(8, 768), torch.empty(8, 768, device=DEV), 8 * 768])
    assert torch.equal(a1[2], a2[2])


@pytest.mark.parametrize("B,H,N,d", [(6, 2, 12, 64), (5, 12, 50, 64), (3, 2, 150, 64), (2, 2, 70, 80)])
def test_key_padded_attention(B, H, N, d):
    torch.manual_seed(N)
    qkv = rnd(B, N, 3 * H * d).to(BF16)
    klen = torch.randint(1, N + 1, (B,), device=DEV, dtype=torch.int32)
    klen[0] = N
    scale = d ** -0.5
    res = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        out = torch.empty(B * N, H * d, device=DEV, dtype=BF16)
        lse = torch.empty(B, H, N, device=DEV)
        fn("attn_padded_fwd", qkv, out, lse, klen, B, N, H, d, scale)
        res.append((out, lse))
    close(res[0][0], res[1][0], atol=2e-2, what="padded attn out")
    close(res[0][1], res[1][1], atol=1e-3, what="padded attn lse")
    out, lse = res[1]
    dout = rnd(B * N, H * d).to(BF16)
    grads = []
    for fn in (L.call, lambda nme, *a: emu.OPS[nme](*a)):
        dqkv = torch.full_like(qkv, float("nan"))
        delta = torch.empty_like(lse)
        fn("attn_padded_bwd", qkv, out, dout, lse, delta, dqkv, klen, B, N, H, d, scale)
        grads.append(dqkv)
    assert torch.isfinite(grads[0].float()).all(), "attn_padded_bwd left elements unwritten"
    close(grads[0], grads[1], atol=3e-2, rtol=3e-2, what="padded attn dqkv")
    g5 = grads[0].view(B, N, 3, H, d)
    for b in range(B):                         # masked keys receive exactly zero dk / dv
        assert (g5[b, int(klen[b]):, 1:] == 0).all()


def test_v1_model_against_reference_golden_and_oracle():
    """TVTS v1 with its DistilBERT text encoder on the CUDA kernels vs the executed-reference fixture and the oracle's gradients."""
    import v1_fixture
    from test_v1_cpu import build
    g, dims, cfg, names, sd, data = v1_fixture.load()
    m = build(dims)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    dd = {"video": data["video"].cuda(), "keep_ind": data["keep_ind"], "label": data["label"].cuda(),
          "text": {k: v.cuda() for k, v in data["text"].items()}}
    te, ve, pred = m(dd)
    l1 = M.NormSoftmaxLoss(0.05)(M.sim_matrix(ve, te))
    l2 = E.sort_ce(pred, dd["label"], 2.0)
    (l1 + l2).backward()
    torch.cuda.synchronize()
    assert abs(l1.item() - float(g["loss1"])) < 5e-2 and abs(l2.item() - float(g["loss2"])) < 5e-2
    np.testing.assert_allclose(te.detach().cpu().numpy(), g["text_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(ve.detach().cpu().numpy(), g["video_emb"], atol=3e-2, rtol=3e-2)
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["pred_order"], atol=5e-2, rtol=5e-2)
    _, _, _, ograds = O.v1_step_with_grads(sd, data["text"], data["video"], data["keep_ind"], data["label"], cfg, dims.heads)
    got = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ograds)
    for k, gr in ograds.items():
        if gr.norm().item() < 1e-6:
            continue
        rel = (got[k].double() - gr.double()).norm().item() / gr.double().norm().item()
        assert rel < 0.08, (k, rel)
