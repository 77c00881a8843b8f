"""Encoder kernels (through the C ABI) against plain fp32 torch references of the same ops."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16)


def close(a, b, rel, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    s = b.abs().max().item()
    assert err <= rel * s + 1e-6, (what, err, s)


@pytest.mark.parametrize("rows,d", [(37, 128), (1000, 768), (513, 1024), (64, 192)])
@pytest.mark.parametrize("with_b", [True, False])
def test_add_layernorm_fwd_bwd(rows, d, with_b):
    from contrastors_b200 import ops
    torch.manual_seed(0)
    a = bf(torch.randn(rows, d, device="cuda"))
    b = bf(torch.randn(rows, d, device="cuda")) if with_b else None
    gamma = (1 + 0.1 * torch.randn(d, device="cuda")).requires_grad_()
    beta = (0.1 * torch.randn(d, device="cuda")).requires_grad_()
    g1 = bf(torch.randn(rows, d, device="cuda"))
    g2 = bf(torch.randn(rows, d, device="cuda"))
    y, stats = ops.add_layernorm_fwd(a, b, gamma.detach(), beta.detach(), 1e-12)
    z = (a.float() + (b.float() if with_b else 0)).requires_grad_()
    ref = F.layer_norm(z, (d,), gamma, beta, 1e-12)
    close(y, ref, 2 ** -7, "y")
    ref.backward(g1.float() + g2.float())
    dgamma = torch.zeros(d, device="cuda")
    dbeta = torch.zeros(d, device="cuda")
    dz = ops.add_layernorm_bwd(a, b, g1, g2, gamma.detach(), stats, dgamma, dbeta)
    close(dz, z.grad, 2 ** -7, "dz")
    close(dgamma, gamma.grad, 2e-3, "dgamma")
    close(dbeta, beta.grad, 2e-3, "dbeta")


def test_embed_layernorm_fwd_bwd():
    from contrastors_b200 import ops
    torch.manual_seed(1)
    V, d, rows = 500, 768, 3000
    word = (0.02 * torch.randn(V, d, device="cuda"))
    typ = (0.02 * torch.randn(2, d, device="cuda"))
    ids = torch.randint(0, V, (rows,), device="cuda")
    gamma = 1 + 0.1 * torch.randn(d, device="cuda")
    beta = 0.1 * torch.randn(d, device="cuda")
    wb, tb = bf(word), bf(typ)
    y, stats = ops.embed_layernorm_fwd(ids, None, wb, tb, gamma, beta, 1e-12)
    w32 = wb.float().requires_grad_()
    t32 = tb.float().requires_grad_()
    g_ = gamma.clone().requires_grad_()
    b_ = beta.clone().requires_grad_()
    ref = F.layer_norm(w32[ids] + t32[0], (d,), g_, b_, 1e-12)
    close(y, ref, 2 ** -7, "y")
    g1 = bf(torch.randn(rows, d, device="cuda"))
    ref.backward(g1.float())
    dword = torch.zeros(V, d, device="cuda")
    dtyp = torch.zeros(2, d, device="cuda")
    dgamma = torch.zeros(d, device="cuda")
    dbeta = torch.zeros(d, device="cuda")
    ops.embed_layernorm_bwd(ids, None, wb, tb, g1, None, gamma, stats, dword, dtyp, dgamma, dbeta)
    close(dword, w32.grad, 2e-3, "dword")
    close(dtyp, t32.grad, 2e-3, "dtype")
    close(dgamma, g_.grad, 2e-3, "dgamma")
    close(dbeta, b_.grad, 2e-3, "dbeta")


def _rope_tables(S, Dh, base):
    inv = 1.0 / (base ** (torch.arange(0, Dh, 2, dtype=torch.float32) / Dh))
    fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
    return torch.cos(fr).cuda(), torch.sin(fr).cuda()


def test_rope_and_positions():
    from contrastors_b200 import ops
    torch.manual_seed(2)
    H, Dh = 3, 64
    lens = [5, 130, 64]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    T = sum(lens)
    pos = ops.token_positions(cu, T)
    want = torch.cat([torch.arange(l) for l in lens]).int().cuda()
    assert torch.equal(pos, want)
    cos_t, sin_t = _rope_tables(256, Dh, 1000.0)
    qkv = bf(torch.randn(T, 3, H, Dh, device="cuda"))
    ref = qkv.float().clone()
    c = cos_t[pos.long()][:, None, :]
    s = sin_t[pos.long()][:, None, :]
    for w in (0, 1):
        x1, x2 = ref[:, w, :, :32].clone(), ref[:, w, :, 32:].clone()
        ref[:, w, :, :32] = x1 * c - x2 * s
        ref[:, w, :, 32:] = x2 * c + x1 * s
    out = qkv.clone().view(T, -1)
    ops.rope_inplace(out, pos, cos_t, sin_t, H, Dh)
    close(out.view(T, 3, H, Dh), ref, 2 ** -7, "rope")
    # transpose rotation undoes it (up to bf16 rounding)
    ops.rope_inplace(out, pos, cos_t, sin_t, H, Dh, backward=True)
    close(out.view(T, 3, H, Dh), qkv, 2 ** -6, "rope^T rope")


def test_swiglu_fwd_bwd():
    from contrastors_b200 import ops
    torch.manual_seed(3)
    T, I = 777, 3072
    yg = bf(torch.randn(T, 2 * I, device="cuda"))
    out = ops.swiglu_fwd(yg)
    x = yg.float().requires_grad_()
    ref = x[:, :I] * F.silu(x[:, I:])
    close(out, ref, 2 ** -7, "swiglu")
    g = bf(torch.randn(T, I, device="cuda"))
    ref.backward(g.float())
    dyg = ops.swiglu_bwd(g, yg)
    close(dyg, x.grad, 2 ** -7, "dswiglu")


def test_mean_pool_and_head():
    from contrastors_b200 import ops
    torch.manual_seed(4)
    d = 768
    lens = [512, 17, 300, 1]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    T = sum(lens)
    h = bf(torch.randn(T, d, device="cuda"))
    pooled = ops.mean_pool_fwd(h, cu)
    ref = torch.stack([h[cu[i]:cu[i + 1]].float().mean(0) for i in range(len(lens))])
    close(pooled, ref, 1e-5, "pool")
    for hamming in (False, True):
        for normalize in (True, False):
            x = ref.clone().requires_grad_()
            e = F.layer_norm(x, (d,)) if hamming else x
            e = e + (e.to(torch.bfloat16).float() - e).detach()  # bf16 cast with straight-through gradient
            e = F.normalize(e, dim=-1) if normalize else e
            out, save = ops.embed_head_fwd(pooled, hamming, normalize)
            close(out, e, 1e-5, "head")
            g = torch.randn(len(lens), d, device="cuda")
            e.backward(g)
            gp = ops.embed_head_bwd(pooled, g, save, hamming, normalize)
            close(gp, x.grad, 1e-4, "head bwd")
    dh = ops.mean_pool_bwd(g, cu, T)
    want = torch.cat([(g[i] / lens[i]).expand(lens[i], d) for i in range(len(lens))])
    close(dh, want, 2 ** -7, "pool bwd")


def _attn_ref(qkv, lens, H, Dh, scale):
    T = qkv.shape[0]
    x = qkv.float().view(T, 3, H, Dh)
    outs, off = [], 0
    for L in lens:
        q, k, v = (x[off:off + L, i].transpose(0, 1) for i in range(3))  # [H, L, Dh]
        s = (q @ k.transpose(1, 2)) * scale
        outs.append((torch.softmax(s, -1) @ v).transpose(0, 1).reshape(L, H * Dh))
        off += L
    return torch.cat(outs, 0)


@pytest.mark.parametrize("lens,H", [([128], 1), ([512, 512], 2), ([300, 17, 512, 129, 1], 3), ([197] * 4, 12), ([640, 1000], 2),
                                    ([64], 1), ([65, 191, 192, 193], 2), ([2048, 77, 1500], 1)])
def test_attention_fwd_bwd(lens, H):
    """Two-threads-per-row forward + transposed-score backward vs fp32 torch (ragged, ViT-like, > 512 and 2048-token lengths:
    the last case walks the backward's three-stage query ring and every barrier phase through 16 query tiles)."""
    from contrastors_b200 import ops
    torch.manual_seed(5)
    Dh = 64
    scale = 1.0 / math.sqrt(Dh)
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    qkv = bf(torch.randn(T, 3 * H * Dh, device="cuda"))
    out, lse = ops.attn_fwd(qkv, cu, max(lens), H, Dh, scale)
    x = qkv.float().requires_grad_()
    ref = _attn_ref(x, lens, H, Dh, scale)
    close(out, ref, 2 ** -6, "attn out")
    # lse check on the first sequence
    L = lens[0]
    xq = qkv.float().view(T, 3, H, Dh)
    s0 = (xq[:L, 0].transpose(0, 1) @ xq[:L, 1].transpose(0, 1).transpose(1, 2)) * scale
    close(lse[:, :L], torch.logsumexp(s0, -1), 1e-3, "lse")
    dout = bf(torch.randn(T, H * Dh, device="cuda"))
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, cu, max(lens), H, Dh, scale)
    g = x.grad.view(T, 3, H, Dh)
    d = dqkv.float().view(T, 3, H, Dh)
    for i, name in enumerate(["dq", "dk", "dv"]):
        close(d[:, i], g[:, i], 2 ** -5, name)


@pytest.mark.parametrize("lens,H", [([512, 512], 2), ([300, 17, 512, 129, 1], 3)])
def test_attention_bwd_rotary_transpose_in_epilogue(lens, H):
    """dk rotated back inside the attention backward's epilogue (position = row index inside the sequence, cos/sin from
    pos * inv_freq) == the separate rotary pass over the dk slot it replaces; dq / dv are untouched by the switch."""
    from contrastors_b200 import ops
    torch.manual_seed(6)
    Dh = 64
    scale = 1.0 / math.sqrt(Dh)
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    pos = ops.token_positions(cu, T)
    inv = (1.0 / (1000.0 ** (torch.arange(0, Dh, 2, dtype=torch.float32) / Dh)))
    fr = torch.outer(torch.arange(max(lens), dtype=torch.float32), inv)
    cos_t, sin_t, inv_freq = torch.cos(fr).cuda(), torch.sin(fr).cuda(), inv.cuda().contiguous()
    qkv = bf(torch.randn(T, 3 * H * Dh, device="cuda"))
    out, lse = ops.attn_fwd(qkv, cu, max(lens), H, Dh, scale)
    dout = bf(torch.randn(T, H * Dh, device="cuda"))
    two = ops.attn_bwd(qkv, out, dout, lse, cu, max(lens), H, Dh, scale, pos, cos_t, sin_t, None).float().view(T, 3, H, Dh)
    one = ops.attn_bwd(qkv, out, dout, lse, cu, max(lens), H, Dh, scale, pos, cos_t, sin_t, inv_freq).float().view(T, 3, H, Dh)
    assert torch.equal(one[:, 2], two[:, 2])
    close(one[:, 0], two[:, 0], 2 ** -7, "dq")   # dq: fp32 atomics in a different order only
    close(one[:, 1], two[:, 1], 2 ** -6, "dk")   # one bf16 rounding instead of two


def test_adamw_and_clip_match_torch():
    from contrastors_b200 import ops
    torch.manual_seed(6)
    n = 100003
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 3
    ref_p = p.clone().requires_grad_()
    opt = torch.optim.AdamW([ref_p], lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        gg = g.clone()
        coef = ops.grad_clip_coef(gg, 1.0)
        ops.adamw_step(p, gg, m, v, shadow, 2e-4, 0.9, 0.999, 1e-8, 0.01, step, grad_scale_dev=coef[1:], zero_grad=True)
        assert torch.allclose(coef[0], g.norm(), rtol=1e-5)
        assert torch.count_nonzero(gg) == 0
    assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-7)
    assert torch.equal(shadow, p.to(torch.bfloat16))


def test_fused_dropout_add_layernorm():
    """dropout_add_layer_norm semantics (layers/block.py:422-431): z = dropout(a)/(1-p) + b, counter-based mask that is a
    pure function of (seed, row, col): same seed -> same mask, backward re-creates it, statistics match p."""
    from contrastors_b200 import ops
    torch.manual_seed(7)
    rows, d, p = 4096, 768, 0.1
    a = bf(torch.randn(rows, d, device="cuda"))
    b = bf(torch.randn(rows, d, device="cuda"))
    gamma = 1 + 0.1 * torch.randn(d, device="cuda")
    beta = 0.1 * torch.randn(d, device="cuda")
    y1, st1, z1 = ops.add_layernorm_fwd(a, b, gamma, beta, 1e-12, want_z=True, p_drop=p, seed=1234)
    y2, st2, z2 = ops.add_layernorm_fwd(a, b, gamma, beta, 1e-12, want_z=True, p_drop=p, seed=1234)
    y3, _, z3 = ops.add_layernorm_fwd(a, b, gamma, beta, 1e-12, want_z=True, p_drop=p, seed=99)
    assert torch.equal(y1, y2) and torch.equal(z1, z2) and not torch.equal(z1, z3)
    # recover the keep mask: z - b is either 0 or a / (1 - p)
    contrib = z1.float() - b.float()
    keep = contrib.abs() > 0.5 * a.float().abs().clamp_min(1e-3) * 0 + 1e-2 * (a.float().abs() > 0.05)
    sel = a.float().abs() > 0.05
    frac = 1.0 - keep[sel].float().mean().item()
    assert abs(frac - p) < 0.01, frac
    # kept entries are scaled by 1/(1-p); dropped ones contribute nothing
    ref_z = torch.where(keep, a.float() / (1 - p), torch.zeros_like(contrib)) + b.float()
    assert (z1.float()[sel] - ref_z[sel]).abs().max().item() <= 0.05
    ref_y = torch.nn.functional.layer_norm(z1.float(), (d,), gamma, beta, 1e-12)
    close(y1, ref_y, 2 ** -6, "y with dropout")
    # backward: dz (for b) is the plain LN backward; da = dz * keep / (1 - p)
    g = bf(torch.randn(rows, d, device="cuda"))
    dz, da = ops.add_layernorm_bwd(a, b, g, None, gamma, st1, None, None, p_drop=p, seed=1234)
    zz = z1.float().requires_grad_()
    torch.nn.functional.layer_norm(zz, (d,), gamma, beta, 1e-12).backward(g.float())
    close(dz, zz.grad, 2 ** -6, "dz")
    want_da = torch.where(keep, dz.float() / (1 - p), torch.zeros_like(dz.float()))
    assert (da.float()[sel] - want_da[sel]).abs().max().item() <= 2 ** -6 * want_da.abs().max().item() + 1e-3


def test_tower_dropout_replays_under_randcontext():
    """GradCache contract (rand_state.py): re-running the tower inside the chunk's RandContext reproduces the embedding
    bit for bit even with dropout on, and without the context the masks differ."""
    import contrastors_b200 as cb
    cfg = cb.NomicBertConfig(vocab_size=256, n_embd=128, n_head=2, n_inner=256, n_layer=2, resid_pdrop=0.1)
    model = cb.BiEncoder(cb.BiEncoderConfig(encoder=cfg)).cuda()
    model.trunk.reset_parameters(seed=1)
    model.train()
    ids = torch.randint(0, 256, (4, 64), device="cuda")
    chunk = {"input_ids": ids}
    torch.manual_seed(5)
    state = cb.RandContext(chunk)
    with torch.no_grad():
        e1 = model(**chunk)["embedding"]
        e_other = model(**chunk)["embedding"]
    assert not torch.equal(e1, e_other)
    with state:
        e2 = model(**chunk)["embedding"]
    assert torch.equal(e1, e2.detach())
    e2.sum().backward()
    assert torch.isfinite(model.trunk.flat_grad()).all() and torch.count_nonzero(model.trunk.flat_grad()) > 0
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(**chunk)["embedding"], model(**chunk)["embedding"])  # no dropout in eval mode
