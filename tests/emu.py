"""TEST INFRASTRUCTURE ONLY -- a torch (CPU or GPU) restatement of every C-ABI op of libtvts_b200.so.

Two uses, both inside tests/:
  * `-m "not gpu"`: `install()` monkeypatches tvts_b200._lib.call / .gemm so the host-side orchestration
    (tvts_b200/engine.py, the model mirror, the trainer) can be checked against the oracle without a GPU;
  * `-m gpu`: the per-kernel parity tests run the real kernel and this restatement on the same inputs.
The product never imports this module (tvts_b200/_lib.py raises if the native library is missing).
"""
import math

import torch

from tvts_b200._lib import OPERAND_DTYPE

BF16, F32 = OPERAND_DTYPE, torch.float32      # the 16-bit operand format of the build under test (bfloat16, or float16 with TVTS_OPERAND=fp16)


def _act(x, a):
    if a in (1, "quick_gelu"):
        return x * torch.sigmoid(1.702 * x)
    if a in (2, "gelu"):
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    return x


def _dact(x, a):
    if a in (1, "quick_gelu"):
        s = torch.sigmoid(1.702 * x)
        return s * (1.0 + 1.702 * x * (1.0 - s))
    if a in (2, "gelu"):
        return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return torch.ones_like(x)


def _mat(t, rows, cols, ld):
    """[rows, cols] matrix at t's data pointer with a row pitch of ld elements (what the C ABI sees: pointer + leading dimension)."""
    return t.as_strided((rows, cols), (ld, 1), t.storage_offset())


def gemm(a, b, out, *, M, N, K, lda, ldb, ldo=None, a_mn=False, b_mn=False, bias=None, residual=None, ldr=None,
         aux=None, ldaux=None, out_pre=None, act=None, dact=None, accumulate=False, splits=0, alpha=1.0):
    ldo = N if ldo is None else ldo
    A = _mat(a, K, M, lda).t() if a_mn else _mat(a, M, K, lda)
    Bm = _mat(b, K, N, ldb).t() if b_mn else _mat(b, N, K, ldb)
    v = alpha * (A.float() @ Bm.float().t())
    if bias is not None:
        v = v + bias.float()
    o = _mat(out, M, N, ldo)
    if out_pre is not None:
        _mat(out_pre, M, N, ldo).copy_(v.to(BF16))
    v = _act(v, act)
    if dact:
        v = v * _dact(_mat(aux, M, N, N if ldaux is None else ldaux).float(), dact)
    if residual is not None:
        v = v + _mat(residual, M, N, N if ldr is None else ldr)
    if accumulate:
        o.add_(v.to(o.dtype))
    else:
        o.copy_(v.to(o.dtype))
    return out


def _sets(mode, N, T, n):
    """list of (stationary tokens, streamed tokens) index lists"""
    if mode == 0:
        return [(list(range(N)), list(range(N)))]
    groups = []
    if mode == 1:
        for f in range(T):
            g = [1 + f * n + i for i in range(n)]
            groups.append((g, [0] + g))
    else:
        for i in range(n):
            g = [1 + f * n + i for f in range(T)]
            groups.append((g, [0] + g))
    groups.append(([0], list(range(N))))
    return groups


def _attn_mask(N, mode, T, n, causal, device):
    m = torch.zeros(N, N, dtype=torch.bool, device=device)
    for q, k in _sets(mode, N, T, n):
        qi = torch.tensor(q, device=device)
        ki = torch.tensor(k, device=device)
        m[qi[:, None], ki[None, :]] = True
    if causal:
        m &= torch.ones(N, N, dtype=torch.bool, device=device).tril()
    return m


def _attn_probs(qkv, B, N, H, d, mode, T, n, causal, scale):
    x = qkv.view(B, N, 3, H, d).float()
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)  # [B,H,N,d]
    s = (q @ k.transpose(-1, -2)) * scale
    mask = _attn_mask(N, mode, T, n, causal, qkv.device)
    s = s.masked_fill(~mask, float("-inf"))
    return q, k, v, s


def attn_fwd(qkv, out, lse, B, N, H, d, mode, T, n, causal, scale):
    q, k, v, s = _attn_probs(qkv, B, N, H, d, mode, T, n, causal, scale)
    l = torch.logsumexp(s, -1)
    p = torch.exp(s - l[..., None])
    o = (p @ v).transpose(1, 2).reshape(B * N, H * d)
    out.view(B * N, H * d).copy_(o.to(BF16))
    lse.view(B, H, N).copy_(l)


def attn_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, d, mode, T, n, causal, scale):
    q, k, v, s = _attn_probs(qkv, B, N, H, d, mode, T, n, causal, scale)
    p = torch.exp(s - lse.view(B, H, N)[..., None])
    do = dout.view(B, N, H, d).float().transpose(1, 2)
    o = out.view(B, N, H, d).float().transpose(1, 2)
    dl = (do * o).sum(-1)
    delta.view(B, H, N).copy_(dl)
    dp = do @ v.transpose(-1, -2)
    ds = p * (dp - dl[..., None])
    dq = (ds @ k) * scale
    dk = (ds.transpose(-1, -2) @ q) * scale
    dv = p.transpose(-1, -2) @ do
    r = torch.stack([dq, dk, dv], 0).permute(1, 3, 0, 2, 4)  # [B,N,3,H,d]
    dqkv.view(B, N, 3, H, d).copy_(r.to(BF16))


def attn_bwd_bias(qkv, out, dout, lse, delta, dqkv, dbias, B, N, H, d, mode, T, n, causal, scale):
    attn_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, d, mode, T, n, causal, scale)
    dbias.add_(dqkv.view(B * N, 3 * H * d).float().sum(0))


def _padded_scores(qkv, klen, B, N, H, d, scale):
    t = qkv.view(B, N, 3, H, d).float()
    q, k, v = (t[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * scale
    ok = torch.arange(N, device=qkv.device)[None, :] < klen.view(B, 1).to(torch.int64)          # [B, N] valid KEYS
    return q, k, v, s.masked_fill(~ok[:, None, None, :], float("-inf"))


def attn_padded_fwd(qkv, out, lse, klen, B, N, H, d, scale):
    q, k, v, s = _padded_scores(qkv, klen, B, N, H, d, scale)
    l = torch.logsumexp(s, -1)
    p = torch.exp(s - l[..., None])
    out.view(B * N, H * d).copy_((p @ v).transpose(1, 2).reshape(B * N, H * d).to(BF16))
    lse.view(B, H, N).copy_(l)


def attn_padded_bwd(qkv, out, dout, lse, delta, dqkv, klen, B, N, H, d, scale):
    q, k, v, s = _padded_scores(qkv, klen, B, N, H, d, scale)
    p = torch.exp(s - lse.view(B, H, N)[..., None])
    do = dout.view(B, N, H, d).float().transpose(1, 2)
    o = out.view(B, N, H, d).float().transpose(1, 2)
    dl = (do * o).sum(-1)
    delta.view(B, H, N).copy_(dl)
    ds = p * (do @ v.transpose(-1, -2) - dl[..., None])
    r = torch.stack([(ds @ k) * scale, (ds.transpose(-1, -2) @ q) * scale, p.transpose(-1, -2) @ do], 0).permute(1, 3, 0, 2, 4)
    dqkv.view(B, N, 3, H, d).copy_(r.to(BF16))


def tubelet_gather(video, keep, cols, B, T, R, p, n):
    g, nt = R // p, T // 2
    x = video.view(B, nt, 2, 3, g, p, g, p).permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(B, nt, g * g, 3 * 2 * p * p)
    idx = keep.view(B, nt, n, 1).expand(B, nt, n, x.shape[-1])
    cols.view(B, nt, n, -1).copy_(torch.gather(x, 2, idx).to(BF16))


def video_assemble_tube(tok, cls, pos, tem, keep, x0, B, nt, n, D):
    pos, tem = pos.view(-1, D), tem.view(-1, D)          # the parameters are [1, P+1, D] / [1, tubes, D]
    t = tok.view(B, nt, n, D) + pos[1:][keep.view(B, nt, n)] + tem[:nt][None, :, None, :]
    x = x0.view(B, 1 + nt * n, D)
    x[:, 0] = cls.reshape(-1) + pos[0]
    x[:, 1:] = t.reshape(B, nt * n, D)


def video_assemble_tube_bwd(dx0, keep, dcls, dpos, dtem, dtok, B, nt, n, D):
    dpos, dtem = dpos.view(-1, D), dtem.view(-1, D)
    d = dx0.view(B, 1 + nt * n, D)
    dcls.view(-1).add_(d[:, 0].sum(0))
    dpos[0].add_(d[:, 0].sum(0))
    dp = d[:, 1:].reshape(B, nt, n, D)
    dtem[:nt].add_(dp.sum((0, 2)))
    dpos.index_add_(0, (1 + keep.view(-1)), dp.reshape(-1, D))
    dtok.view(B * nt * n, D).copy_(dp.reshape(-1, D).to(BF16))


def relu_bf16(x, y, n):
    y.view(-1).copy_(torch.relu(x.reshape(-1)).to(BF16))


def relu_bwd(x, dy, dx, n):
    dx.view(-1).copy_(dy.reshape(-1) * (x.reshape(-1) > 0))


def attn_window_fwd(qkv, out, lse, B, N, H, d, q0, qn, scale):
    o = torch.empty(B * N, H * d, dtype=BF16, device=qkv.device)
    l = torch.empty(B, H, N, device=qkv.device)
    attn_fwd(qkv, o, l, B, N, H, d, 0, 0, 0, 0, scale)
    out.view(B, N, H * d)[:, q0:q0 + qn].copy_(o.view(B, N, H * d)[:, q0:q0 + qn])
    lse.view(B, H, N)[:, :, q0:q0 + qn].copy_(l[:, :, q0:q0 + qn])


def attn_window_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, d, q0, qn, scale):
    # queries outside the window do not exist: mask their probabilities by zeroing dout / out there and using a huge lse
    o = torch.zeros(B, N, H * d, dtype=out.dtype, device=out.device)
    do = torch.zeros_like(o)
    l = torch.full((B, H, N), 1e30, device=lse.device)
    o[:, q0:q0 + qn] = out.view(B, N, H * d)[:, q0:q0 + qn]
    do[:, q0:q0 + qn] = dout.view(B, N, H * d)[:, q0:q0 + qn]
    l[:, :, q0:q0 + qn] = lse.view(B, H, N)[:, :, q0:q0 + qn]
    attn_bwd(qkv, o.view(B * N, H * d), do.view(B * N, H * d), l, delta, dqkv, B, N, H, d, 0, 0, 0, 0, scale)


def layernorm_fwd(x, g, b, y, y_bf16, mean, rstd, M, D, eps):
    x = x.view(M, D)
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    rs = torch.rsqrt((xc * xc).mean(-1, keepdim=True) + eps)
    y.view(M, D).copy_((xc * rs * g + b).to(y.dtype))
    if mean is not None:
        mean.copy_(mu.view(-1))
    if rstd is not None:
        rstd.copy_(rs.view(-1))


def layernorm_bwd(dy, dy_bf16, x, mean, rstd, g, res1, res2, dx, dxb, dg, db, M, D):
    d = dy.view(M, D).float()
    xh = (x.view(M, D) - mean[:, None]) * rstd[:, None]
    gy = d * g
    s1 = gy.mean(-1, keepdim=True)
    s2 = (gy * xh).mean(-1, keepdim=True)
    o = rstd[:, None] * (gy - s1 - xh * s2)
    if res1 is not None:
        o = o + res1.view(M, D)
    if res2 is not None:
        o = o + res2.view(M, D)
    if dx is not None:
        dx.view(M, D).copy_(o)
    if dxb is not None:
        dxb.view(M, D).copy_(o.to(BF16))
    if dg is not None:
        dg.add_((d * xh).sum(0))
        db.add_(d.sum(0))


def layernorm_bwd_colsum(dy, dy_bf16, x, mean, rstd, g, res1, res2, dx, dxb, dg, db, dxsum, M, D):
    tmp = dx if dx is not None else torch.empty(M, D, dtype=torch.float32, device=x.device)
    layernorm_bwd(dy, dy_bf16, x, mean, rstd, g, res1, res2, tmp, dxb, dg, db, M, D)
    if dxsum is not None:
        dxsum += tmp.view(M, D).sum(0)


def cast_bf16(src, dst, n):
    dst.view(-1).copy_(src.reshape(-1).to(BF16))


def colsum_bf16(x, out, M, N, ld):
    out.add_(_mat(x, M, N, ld).float().sum(0))


def patch_gather(video, keep, cols, B, T, R, p, n):
    g = R // p
    v = video.view(B, T, 3, g, p, g, p).permute(0, 1, 3, 5, 2, 4, 6).reshape(B, T, g * g, 3 * p * p)
    idx = keep.view(B, 1, n, 1).expand(B, T, n, 3 * p * p)
    cols.view(B, T, n, -1).copy_(torch.gather(v, 2, idx).to(BF16))


def patch_gather_u8(video, keep, cols, B, T, R, p, n, mean, std):
    m = torch.tensor(list(mean), dtype=torch.float32, device=video.device)[None, None, :, None, None]
    s = torch.tensor(list(std), dtype=torch.float32, device=video.device)[None, None, :, None, None]
    x = video.view(B, T, 3, R, R).float().div(255).sub_(m).div_(s)        # ClipToTensor + Normalize of the reference, in its order
    patch_gather(x, keep, cols, B, T, R, p, n)


def patch_gather_ld(video, keep, cols, B, T, R, p, n, ld):
    K = 3 * p * p
    tmp = torch.empty(B * T * n, K, dtype=BF16, device=video.device)
    patch_gather(video, keep, tmp, B, T, R, p, n)
    c = cols.view(B * T * n, ld)
    c.zero_()
    c[:, :K].copy_(tmp)


def cast_bf16_pad(src, dst, rows, cols, ld):
    d = dst.view(rows, ld)
    d.zero_()
    d[:, :cols].copy_(src.reshape(rows, cols).to(BF16))


def attn_generic_fwd(*a):
    attn_fwd(*a)


def attn_generic_bwd(*a):
    attn_bwd(*a)


def video_assemble(tok, cls, pos, tem, keep, x0, B, T, n, D):
    t = tok.view(B, T, n, D) + pos[1:][keep.view(B, n)][:, None] + tem[:T][None, :, None, :]
    x = x0.view(B, 1 + T * n, D)
    x[:, 0] = cls + pos[0]
    x[:, 1:] = t.reshape(B, T * n, D)


def video_assemble_bwd(dx0, keep, dcls, dpos, dtem, dtok, B, T, n, D):
    d = dx0.view(B, 1 + T * n, D)
    dcls.add_(d[:, 0].sum(0))
    dpos[0].add_(d[:, 0].sum(0))
    dt = d[:, 1:].reshape(B, T, n, D)
    dtem[:T].add_(dt.sum((0, 2)))
    dpos.index_add_(0, (1 + keep.view(B, n)).reshape(-1), dt.sum(1).reshape(B * n, D))
    dtok.view(B, T, n, D).copy_(dt.to(BF16))


def text_embed(tok, is64, table, pos, x, rows, L, W):
    x.view(rows, L, W).copy_(table[tok.view(rows, L).long()] + pos.view(-1, W)[:L])      # the kernel reads rows 0..L-1 of the position table


def text_embed_bwd(dx, tok, is64, dtable, dpos, rows, L, W):
    d = dx.view(rows, L, W)
    if dpos is not None:
        dpos.view(-1, W)[:L].add_(d.sum(0))
    if dtable is not None:
        dtable.index_add_(0, tok.view(-1).long(), d.reshape(rows * L, W))


def argmax_rows(tok, is64, flat_idx, rows, L):
    flat_idx.copy_(torch.arange(rows, device=tok.device) * L + tok.view(rows, L).long().argmax(-1))


def gather_rows(src, idx, dst, rows, D):
    dst.view(rows, D).copy_(src.view(-1, D)[idx])


def scatter_rows(src, idx, dst, rows, D, accumulate):
    if accumulate:
        dst.view(-1, D).index_add_(0, idx, src.view(rows, D))
    else:
        dst.view(-1, D)[idx] = src.view(rows, D)


def group_mean(t, out, nt, B, E):
    out.view(B, E).copy_(t.view(nt, B, E).mean(0))


def group_mean_bwd(dout, dt, dt_bf16, nt, B, E):
    v = (dout.view(1, B, E) / nt).expand(nt, B, E)
    if dt is not None:
        dt.view(nt, B, E).copy_(v)
    if dt_bf16 is not None:
        dt_bf16.view(nt, B, E).copy_(v.to(BF16))


def sort_concat(vtok, t, te, z, B, N, nt, E):
    zz = z.view(B, N + nt, E)
    te = te.view(2, E)
    zz[:, :N] = vtok.view(B, N, E) + te[0]
    zz[:, N:] = t.view(nt, B, E).permute(1, 0, 2) + te[1]


def sort_concat_bwd(dz, dvtok, dte, B, N, nt, E):
    d = dz.view(B, N + nt, E)
    dvtok.view(B, N, E).copy_(d[:, :N])
    dte.view(2, E)[0].add_(d[:, :N].sum((0, 1)))
    dte.view(2, E)[1].add_(d[:, N:].sum((0, 1)))


def add_rows(src, dst, rows, D, lds, ldd):
    dst.view(-1)[: rows * ldd].view(rows, ldd)[:, :D].add_(src.view(-1)[: rows * lds].view(rows, lds)[:, :D])


def small_linear_fwd(x, w, bias, y, R, K, O):
    y.view(R, O).copy_(x.view(R, K) @ w.view(O, K).t() + (bias if bias is not None else 0))


def small_linear_bwd(dy, x, w, dx, dw, db, R, K, O):
    dy = dy.view(R, O)
    dx.view(R, K).copy_(dy @ w.view(O, K))
    dw.view(O, K).add_(dy.t() @ x.view(R, K))
    if db is not None:
        db.add_(dy.sum(0))


def normalize_rows(x, xn, norm, rows, E, eps):
    nr = x.view(rows, E).norm(dim=1)
    norm.copy_(nr)
    xn.view(rows, E).copy_(x.view(rows, E) / nr.clamp_min(eps)[:, None])


def sim_matrix(an, bn, S, Ra, Rb, E, scale):
    S.view(Ra, Rb).copy_(scale * an.view(Ra, E) @ bn.view(Rb, E).t())


def sim_matrix_bwd(G, self_n, other_n, self_norm, dself, R_self, R_other, E, row0, nrows, transposed, scale, eps):
    Gm = G.view(R_other, R_self).t() if transposed else G.view(R_self, R_other)
    dn = scale * Gm[row0:row0 + nrows] @ other_n.view(R_other, E)
    sn = self_n.view(R_self, E)[row0:row0 + nrows]
    nr = self_norm[row0:row0 + nrows]
    proj = (dn - sn * (sn * dn).sum(-1, keepdim=True)) / nr.clamp_min(eps)[:, None]
    clamped = (nr <= eps)[:, None]
    dself.view(nrows, E).copy_(torch.where(clamped, dn / eps, proj))


def nsl_fwd(S, lse_r, lse_c, loss, Bg, temperature):
    x = S.view(Bg, Bg) / temperature
    lr, lc = torch.logsumexp(x, 1), torch.logsumexp(x, 0)
    lse_r.copy_(lr)
    lse_c.copy_(lc)
    d = torch.diagonal(x)
    loss.copy_(-(d - lr).mean() - (d - lc).mean())


def nsl_bwd(S, lse_r, lse_c, gout, G, Bg, temperature):
    x = S.view(Bg, Bg) / temperature
    g = (gout if gout is not None else 1.0) / temperature / Bg
    G.view(Bg, Bg).copy_((torch.exp(x - lse_r[:, None]) + torch.exp(x - lse_c[None, :]) - 2 * torch.eye(Bg, device=S.device)) * g)


def sort_ce(logits, labels, gout, loss, dlogits, R, C, weight):
    x = logits.view(R, C)
    l = torch.logsumexp(x, 1)
    if loss is not None:
        loss.copy_(weight * (l - x[torch.arange(R, device=x.device), labels]).mean())
    if dlogits is not None:
        g = (gout if gout is not None else 1.0) * weight / R
        oh = torch.zeros_like(x)
        oh[torch.arange(R, device=x.device), labels] = 1.0
        dlogits.view(R, C).copy_((torch.exp(x - l[:, None]) - oh) * g)


def contrastive_sortce_fused(video_all, text_all, Bg, E, row0, nloc, temperature, eps, logits, labels, R, C, weight, loss1, loss2, dv, dt, dx):
    a = video_all.detach().clone().float().requires_grad_(True)
    b = text_all.detach().clone().float().requires_grad_(True)
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    z = an @ bn.t() / temperature
    l1 = -torch.diagonal(torch.log_softmax(z, 1)).mean() - torch.diagonal(torch.log_softmax(z.t(), 1)).mean()
    l1.backward()
    loss1.copy_(l1.detach())
    dv.copy_(a.grad[row0:row0 + nloc])
    dt.copy_(b.grad[row0:row0 + nloc])
    if logits is not None:
        x = logits.detach().clone().float().requires_grad_(True)
        l2 = weight * torch.nn.functional.cross_entropy(x.view(R, C), labels.view(R).long())
        l2.backward()
        loss2.copy_(l2.detach())
        if dx is not None:
            dx.copy_(x.grad)


def adamw_flat(p, g, m, v, pb, chunk_tensor, table, n_chunks, chunk, b1, b2, eps, gscale):
    t = table[chunk_tensor.long()]                                   # [n_chunks, 4]
    step = t[:, 0:1]; lrwd = t[:, 1:2]; act = t[:, 2:3] != 0
    P, G, Mm, V = (x.view(n_chunks, chunk) for x in (p, g, m, v))
    gg = G * gscale
    m2 = Mm * b1 + (1.0 - b1) * gg
    v2 = V * b2 + (1.0 - b2) * gg * gg
    x = P - step * (m2 / (v2.sqrt() + eps))
    x = torch.where(lrwd > 0, x - lrwd * x, x)
    Mm.copy_(torch.where(act, m2, Mm)); V.copy_(torch.where(act, v2, V)); P.copy_(torch.where(act, x, P))
    if pb is not None:
        PB = pb.view(n_chunks, chunk)
        PB.copy_(torch.where(act, P.to(BF16), PB))


def adamw_flat_dyn(p, g, m, v, pb, chunk_tensor, table, steps, state, n_tensors, n_chunks, chunk, b1, b2, eps, growth_interval, max_scale):
    bad = not bool(torch.isfinite(g).all())
    if not bad:
        t = (steps.float() + 1.0)
        tab = table.clone()
        corr = torch.sqrt(1.0 - torch.pow(torch.tensor(b2), t)) / (1.0 - torch.pow(torch.tensor(b1), t))
        tab[:, 0] = torch.where(table[:, 3] != 0, table[:, 0] * corr, table[:, 0])
        adamw_flat(p, g, m, v, pb, chunk_tensor, tab, n_chunks, chunk, b1, b2, eps, 1.0 / float(state[0]))
        steps += (table[:, 2] != 0).to(steps.dtype)
        state[1] += 1
        if float(state[1]) >= growth_interval:
            state[0] = min(float(state[0]) * 2.0, max_scale)
            state[1] = 0
    else:
        state[0] = max(float(state[0]) * 0.5, 1.0)
        state[1] = 0
        state[3] += 1
    state[2] = 0


def adamw_dyn_check(g, n_elems, state):
    if not bool(torch.isfinite(g.view(-1)[:n_elems]).all()):
        state[2] = 1.0


def adamw_dyn_apply(p, g, m, v, pb, chunk_tensor, table, steps, state, n_chunks, chunk, b1, b2, eps):
    if float(state[2]) != 0.0:
        return
    t = (steps.float() + 1.0)
    tab = table.clone()
    corr = torch.sqrt(1.0 - torch.pow(torch.tensor(b2), t)) / (1.0 - torch.pow(torch.tensor(b1), t))
    tab[:, 0] = torch.where(table[:, 3] != 0, table[:, 0] * corr, table[:, 0])
    n = n_chunks * chunk
    adamw_flat(p.view(-1)[:n], g.view(-1)[:n], m.view(-1)[:n], v.view(-1)[:n], None if pb is None else pb.view(-1)[:n], chunk_tensor[:n_chunks], tab,
               n_chunks, chunk, b1, b2, eps, 1.0 / float(state[0]))


def adamw_dyn_finish(steps, table, state, n_tensors, growth_interval, max_scale):
    if float(state[2]) == 0.0:
        steps += (table[:, 2] != 0).to(steps.dtype)
        state[1] += 1
        if float(state[1]) >= growth_interval:
            state[0] = min(float(state[0]) * 2.0, max_scale)
            state[1] = 0
    else:
        state[0] = max(float(state[0]) * 0.5, 1.0)
        state[1] = 0
        state[3] += 1
    state[2] = 0


OPS = {k: v for k, v in list(globals().items()) if callable(v) and not k.startswith("_") and k not in ("install", "uninstall", "gemm")}

_saved = {}


def install():
    """Route tvts_b200._lib.call / gemm to the restatement above (CPU tests of the host logic)."""
    from tvts_b200 import _lib as L
    if _saved:
        return
    _saved["call"], _saved["gemm"] = L.call, L.gemm
    L.call = lambda name, *args: OPS[name](*args)
    L.gemm = gemm


def uninstall():
    from tvts_b200 import _lib as L
    if _saved:
        L.call, L.gemm = _saved.pop("call"), _saved.pop("gemm")
