"""Tensor-level wrappers over the C ABI (pointer plumbing only; PyTorch is used for device memory and streams)."""
from __future__ import annotations

import torch

from . import _lib

MAJOR_K, MAJOR_MN = 0, 1
_DT = {torch.bfloat16: 0, torch.float32: 1}


class KernelTimer:
    """CUDA-event timing of individual launches on the launching stream (bench.py's live roofline numbers).
    ``sample_every`` = record one launch in N per kind, so the timed region is not perturbed."""

    def __init__(self, sample_every=1):
        self.sample_every = max(1, int(sample_every))
        self.records = []  # (kind, work, start_event, end_event)
        self.counts = {}

    def begin(self, kind):
        c = self.counts.get(kind, 0)
        self.counts[kind] = c + 1
        if c % self.sample_every:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def end(self, kind, work, ev):
        if ev is None:
            return
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record()
        self.records.append((kind, work, ev, e2))

    def summary(self):
        """{kind: dict(launches, sampled, ms_avg, work_avg)} -- call after a device synchronize."""
        out = {}
        for kind, work, a, b in self.records:
            d = out.setdefault(kind, dict(sampled=0, ms=0.0, work=0.0))
            d["sampled"] += 1
            d["ms"] += a.elapsed_time(b)
            d["work"] += work
        for kind, d in out.items():
            d["launches"] = self.counts.get(kind, 0)
            d["ms_avg"] = d["ms"] / d["sampled"]
            d["work_avg"] = d["work"] / d["sampled"]
        return out


TIMER = None  # set to a KernelTimer to time launches


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("contrastors_b200 ops run on a B200 only: got a CPU tensor (no CPU fallback exists)")


def gemm_select_cluster(max_cluster_ctas=0):
    """A/B switch of the GEMM launch mode (0 = default, 1 = single CTA, 2 = CTA pairs, 4 = pairs sharing B by multicast)."""
    _lib.check(_lib.load().cx_gemm_select_cluster(int(max_cluster_ctas)), "cx_gemm_select_cluster")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_major=MAJOR_K, b_major=MAJOR_K, out=None, out_dtype=torch.bfloat16,
         accumulate=False, alpha=1.0):
    """C[M,N] (+)= alpha * A (x) B on tcgen05.  a: [M,K] (K-major) or [K,M] (MN-major); b: [N,K] or [K,N]."""
    _require_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[0], a.shape[1]) if a_major == MAJOR_K else (a.shape[1], a.shape[0])
    N, Kb = (b.shape[0], b.shape[1]) if b_major == MAJOR_K else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape)
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    lib = _lib.load()
    ev = TIMER.begin("gemm") if TIMER is not None else None
    _lib.check(lib.cx_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a_major, b_major, a.stride(0),
                                b.stride(0), out.stride(0), _DT[out.dtype], int(accumulate), float(alpha), _stream()),
               "cx_gemm_bf16")
    if ev is not None:
        TIMER.end("gemm", 2.0 * M * N * K, ev)
    return out


def gemm_qkv_rope(x, w, pos, inv_freq, rope_cols):
    """qkv = x w^T with RoPE applied to the q and k heads in the GEMM epilogue (inv_freq: fp32 [32])."""
    _require_cuda(x, w)
    T, K = x.shape
    n_out = w.shape[0]
    out = torch.empty(T, n_out, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    ev = TIMER.begin("gemm") if TIMER is not None else None
    _lib.check(lib.cx_gemm_qkv_rope(x.data_ptr(), w.data_ptr(), out.data_ptr(), T, n_out, K, x.stride(0), w.stride(0),
                                    out.stride(0), pos.data_ptr(), inv_freq.data_ptr(), rope_cols, _stream()),
               "cx_gemm_qkv_rope")
    if ev is not None:
        TIMER.end("gemm", 2.0 * T * n_out * K, ev)
    return out


def gemm_swiglu(x, w1, keep_preact=True):
    """(x fc11^T) * silu(x fc12^T) with w1 = [fc11; fc12]; returns (act [M,I], yg [M,2I] or None)."""
    _require_cuda(x, w1)
    M, K = x.shape
    I = w1.shape[0] // 2
    act = torch.empty(M, I, device=x.device, dtype=torch.bfloat16)
    yg = torch.empty(M, 2 * I, device=x.device, dtype=torch.bfloat16) if keep_preact else None
    lib = _lib.load()
    ev = TIMER.begin("gemm") if TIMER is not None else None
    _lib.check(lib.cx_gemm_swiglu(x.data_ptr(), w1.data_ptr(), act.data_ptr(), _ptr(yg), M, I, K, x.stride(0), w1.stride(0),
                                  act.stride(0), 2 * I, _stream()), "cx_gemm_swiglu")
    if ev is not None:
        TIMER.end("gemm", 2.0 * M * 2 * I * K, ev)
    return act, yg


def gemm_swiglu_bwd(dout, w2, yg):
    """dyg [M, 2I] = gradient of [y | gate] given dout [M, d] (gradient of the MLP output) and fc2.weight w2 [d, I]: the fc2 input
    gradient stays in TMEM and the SwiGLU backward runs in the GEMM epilogue."""
    _require_cuda(dout, w2, yg)
    M, K = dout.shape
    I = w2.shape[1]
    assert w2.shape[0] == K and yg.shape == (M, 2 * I)
    dyg = torch.empty_like(yg)
    lib = _lib.load()
    ev = TIMER.begin("gemm") if TIMER is not None else None
    _lib.check(lib.cx_gemm_swiglu_bwd(dout.data_ptr(), w2.data_ptr(), yg.data_ptr(), dyg.data_ptr(), M, I, K, dout.stride(0),
                                      w2.stride(0), yg.stride(0), dyg.stride(0), _stream()), "cx_gemm_swiglu_bwd")
    if ev is not None:
        TIMER.end("gemm", 2.0 * M * I * K, ev)
    return dyg


def rows_to_bf16(x: torch.Tensor, k=None, normalize=False, want_inv_norm=False):
    """bf16 copy of x[:, :k] (optionally L2-normalised) and the per-row inverse norms."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, full = x.shape
    k = full if k is None else k
    ldy = (k + 7) // 8 * 8
    y = torch.empty(rows, ldy, device=x.device, dtype=torch.bfloat16)
    if ldy != k:
        y.zero_()
    inv = torch.empty(rows, device=x.device, dtype=torch.float32) if (want_inv_norm or normalize) else None
    lib = _lib.load()
    _lib.check(lib.cx_rows_to_bf16(x.data_ptr(), x.stride(0), y.data_ptr(), ldy, _ptr(inv), rows, k, int(normalize),
                                   _stream()), "cx_rows_to_bf16")
    return y, inv


def rows_to_bf16_into(x: torch.Tensor, y: torch.Tensor):
    """bf16 copy of fp32 rows into an existing (possibly strided-row) bf16 buffer."""
    _require_cuda(x, y)
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.stride(1) == 1 and y.stride(1) == 1
    lib = _lib.load()
    _lib.check(lib.cx_rows_to_bf16(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), 0, x.shape[0], x.shape[1], 0,
                                   _stream()), "cx_rows_to_bf16")
    return y


def row_inv_norms(x: torch.Tensor, k: int):
    _require_cuda(x)
    inv = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.cx_rows_to_bf16(x.data_ptr(), x.stride(0), 0, 0, inv.data_ptr(), x.shape[0], k, 1, _stream()),
               "cx_rows_to_bf16")
    return inv


def l2norm_bwd(x, g, inv_norm, k, out=None, g_prescaled=False, accumulate=False):
    _require_cuda(x, g, inv_norm)
    rows = x.shape[0]
    if out is None:
        out = torch.zeros_like(x) if k != x.shape[1] else torch.empty_like(x)
        accumulate = False
    assert g.stride(1) == 1 and out.stride(1) == 1 and x.stride(1) == 1
    lib = _lib.load()
    _lib.check(lib.cx_l2norm_bwd(x.data_ptr(), x.stride(0), g.data_ptr(), g.stride(0), inv_norm.data_ptr(),
                                 out.data_ptr(), out.stride(0), rows, k, int(g_prescaled), int(accumulate), _stream()),
               "cx_l2norm_bwd")
    return out


def infonce_workspace(n, m, k_dim, device):
    nbytes = _lib.load().cx_infonce_workspace_bytes(n, m, k_dim)
    return torch.empty(nbytes, device=device, dtype=torch.uint8)


def infonce_fwd(q, d, k_dim, scale, scale_dev, rq, rd, label_offset, label_stride, workspace):
    """Returns (lse[n] f32, argmax[n] i32, label_logit[n] f32, stats[4] f32)."""
    _require_cuda(q, d)
    n, m = q.shape[0], d.shape[0]
    dev = q.device
    lse = torch.empty(n, device=dev, dtype=torch.float32)
    argmax = torch.empty(n, device=dev, dtype=torch.int32)
    label_logit = torch.empty(n, device=dev, dtype=torch.float32)
    stats = torch.zeros(4, device=dev, dtype=torch.float32)
    lib = _lib.load()
    ev = TIMER.begin("infonce_fwd") if TIMER is not None else None
    _lib.check(lib.cx_infonce_fwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, k_dim, float(scale),
                                  _ptr(scale_dev), _ptr(rq), _ptr(rd), label_offset, label_stride, lse.data_ptr(),
                                  argmax.data_ptr(), label_logit.data_ptr(), stats.data_ptr(), workspace.data_ptr(),
                                  _stream()), "cx_infonce_fwd")
    if ev is not None:
        TIMER.end("infonce_fwd", 2.0 * n * m * k_dim, ev)
    return lse, argmax, label_logit, stats


def infonce_mat_fwd(q, d, dims, scale, scale_dev, rq, rd, label_offset, label_stride):
    """Matryoshka statistics for every prefix in ``dims`` (ascending multiples of 64) from ONE accumulation over K.
    rq [P, n], rd [P, m] fp32.  Returns (lse [P, n], argmax [P, n] i32, label_logit [P, n], stats [P, 4])."""
    import ctypes
    _require_cuda(q, d, rq, rd)
    n, m, P = q.shape[0], d.shape[0], len(dims)
    dev = q.device
    lse = torch.empty(P, n, device=dev, dtype=torch.float32)
    argmax = torch.empty(P, n, device=dev, dtype=torch.int32)
    label_logit = torch.empty(P, n, device=dev, dtype=torch.float32)
    stats = torch.empty(P, 4, device=dev, dtype=torch.float32)
    lib = _lib.load()
    ws = torch.empty(lib.cx_infonce_mat_workspace_bytes(n, m, P), device=dev, dtype=torch.uint8)
    dims_c = (ctypes.c_int32 * P)(*[int(k) for k in dims])
    ev = TIMER.begin("infonce_fwd") if TIMER is not None else None
    _lib.check(lib.cx_infonce_mat_fwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, P, ctypes.addressof(dims_c), float(scale),
                                      _ptr(scale_dev), rq.data_ptr(), rd.data_ptr(), label_offset, label_stride, lse.data_ptr(),
                                      argmax.data_ptr(), label_logit.data_ptr(), stats.data_ptr(), ws.data_ptr(), _stream()),
               "cx_infonce_mat_fwd")
    if ev is not None:
        TIMER.end("infonce_fwd", 2.0 * n * m * max(dims), ev)
    return lse, argmax, label_logit, stats


def infonce_mat_bwd(q, d, dims, wrel, scale, scale_dev, rq, rd, label_offset, label_stride, lse, coef, coef_gamma_dev, inv_gamma_dev,
                    width):
    """Single-accumulation Matryoshka backward (2..4 ascending dims).  Returns (dq_raw [n, width], dd_raw [m, width] fp32 -- the
    scale * T_t * operand part of the gradients, zero beyond max(dims) --, alpha [P, n], beta [P, m])."""
    import ctypes
    n, m, P, K = q.shape[0], d.shape[0], len(dims), max(dims)
    dev = q.device
    ldw = (width + 3) // 4 * 4
    mk = (torch.zeros if width > K else torch.empty)
    dq = mk(n, ldw, device=dev, dtype=torch.float32)[:, :width]
    dd = mk(m, ldw, device=dev, dtype=torch.float32)[:, :width]
    alpha = torch.empty(P, n, device=dev, dtype=torch.float32)
    beta = torch.empty(P, m, device=dev, dtype=torch.float32)
    lib = _lib.load()
    ws = torch.empty(lib.cx_infonce_mat_bwd_workspace_bytes(n, m, K, P), device=dev, dtype=torch.uint8)
    dims_c = (ctypes.c_int32 * P)(*[int(k) for k in dims])
    wrel_c = (ctypes.c_float * P)(*[float(w) for w in wrel])
    ev = TIMER.begin("infonce_bwd") if TIMER is not None else None
    _lib.check(lib.cx_infonce_mat_bwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, P, ctypes.addressof(dims_c),
                                      ctypes.addressof(wrel_c), float(scale), _ptr(scale_dev), rq.data_ptr(), rd.data_ptr(), label_offset,
                                      label_stride, lse.data_ptr(), float(coef), coef_gamma_dev.data_ptr(), inv_gamma_dev.data_ptr(),
                                      dq.data_ptr(), dq.stride(0), dd.data_ptr(), dd.stride(0), alpha.data_ptr(), beta.data_ptr(),
                                      ws.data_ptr(), _stream()), "cx_infonce_mat_bwd")
    if ev is not None:
        TIMER.end("infonce_bwd", 4.0 * n * m * K, ev)
    return dq, dd, alpha, beta


def infonce_bwd(q, d, k_dim, scale, scale_dev, rq, rd, label_offset, label_stride, lse, coef, coef_dev, dq, dd,
                accumulate_dd, stats, workspace):
    lib = _lib.load()
    n, m = q.shape[0], d.shape[0]
    ev = TIMER.begin("infonce_bwd") if TIMER is not None else None
    _lib.check(lib.cx_infonce_bwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, k_dim, float(scale),
                                  _ptr(scale_dev), _ptr(rq), _ptr(rd), label_offset, label_stride, lse.data_ptr(),
                                  float(coef), _ptr(coef_dev), dq.data_ptr(), dq.stride(0), dd.data_ptr(), dd.stride(0),
                                  int(accumulate_dd), stats.data_ptr(), workspace.data_ptr(), _stream()),
               "cx_infonce_bwd")
    if ev is not None:
        TIMER.end("infonce_bwd", 4.0 * n * m * k_dim, ev)


# ------------------------------------------------------------------------------------------------ encoder ops
def _ws_bytes(nbytes, device):
    return torch.empty(max(int(nbytes), 16), device=device, dtype=torch.uint8)


def add_layernorm_fwd(a, b, gamma, beta, eps, want_z=False, p_drop=0.0, seed=0):
    """y = LN(dropout(a) + b) (b may be None); returns (y bf16 [rows,d], stats fp32 [rows,2]) (+ z when want_z)."""
    _require_cuda(a, b)
    rows, d = a.shape
    y = torch.empty_like(a)
    z = torch.empty_like(a) if want_z else None
    stats = torch.empty(rows, 2, device=a.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.cx_add_layernorm_fwd(a.data_ptr(), _ptr(b), _ptr(gamma), _ptr(beta), y.data_ptr(), stats.data_ptr(),
                                        rows, d, float(eps), _ptr(z), float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
               "cx_add_layernorm_fwd")
    return (y, stats, z) if want_z else (y, stats)


def add_layernorm_bwd(a, b, g1, g2, gamma, stats, dgamma, dbeta, gres=None, p_drop=0.0, seed=0):
    """dz (bf16) for z = dropout(a) + b given upstream g1 (+ g2) (+ gres on the residual stream); ADDS into dgamma/dbeta.
    With dropout returns (dz, da): dz is the gradient of b / the residual, da that of the dropped branch a."""
    rows, d = a.shape
    dz = torch.empty_like(a)
    da = torch.empty_like(a) if p_drop > 0 else None
    lib = _lib.load()
    ws = _ws_bytes(lib.cx_layernorm_bwd_workspace_bytes(d), a.device) if dgamma is not None else None
    _lib.check(lib.cx_add_layernorm_bwd(a.data_ptr(), _ptr(b), g1.data_ptr(), _ptr(g2), _ptr(gamma), stats.data_ptr(),
                                        dz.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(ws), rows, d, _ptr(gres), float(p_drop),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(da), _stream()), "cx_add_layernorm_bwd")
    return (dz, da) if p_drop > 0 else dz


def embed_layernorm_fwd(ids, type_ids, word_emb, type_emb, gamma, beta, eps, p_drop=0.0, seed=0):
    _require_cuda(ids, word_emb)
    rows = ids.numel()
    d = word_emb.shape[1]
    y = torch.empty(rows, d, device=ids.device, dtype=torch.bfloat16)
    stats = torch.empty(rows, 2, device=ids.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.cx_embed_layernorm_fwd(ids.data_ptr(), _ptr(type_ids), word_emb.data_ptr(), type_emb.data_ptr(),
                                          gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), stats.data_ptr(), rows, d,
                                          float(eps), float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
               "cx_embed_layernorm_fwd")
    return y, stats


def embed_layernorm_bwd(ids, type_ids, word_emb, type_emb, g1, g2, gamma, stats, dword, dtype_emb, dgamma, dbeta,
                        padding_idx=-1, p_drop=0.0, seed=0):
    rows = ids.numel()
    d = word_emb.shape[1]
    lib = _lib.load()
    ws = _ws_bytes(lib.cx_layernorm_bwd_workspace_bytes(d), ids.device)
    _lib.check(lib.cx_embed_layernorm_bwd(ids.data_ptr(), _ptr(type_ids), word_emb.data_ptr(), type_emb.data_ptr(),
                                          g1.data_ptr(), _ptr(g2), gamma.data_ptr(), stats.data_ptr(), dword.data_ptr(),
                                          dtype_emb.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), rows, d,
                                          int(padding_idx), float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
               "cx_embed_layernorm_bwd")


def token_positions(cu_seqlens, total_tokens):
    nseq = cu_seqlens.numel() - 1
    pos = torch.empty(total_tokens, device=cu_seqlens.device, dtype=torch.int32)
    lib = _lib.load()
    _lib.check(lib.cx_token_positions(cu_seqlens.data_ptr(), nseq, pos.data_ptr(), 0, _stream()), "cx_token_positions")
    return pos


def rope_inplace(qkv, pos, cos_t, sin_t, H, Dh, backward=False, first_slot=0, num_slots=2):
    T = qkv.shape[0]
    lib = _lib.load()
    _lib.check(lib.cx_rope_inplace(qkv.data_ptr(), pos.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), T, H, Dh,
                                   int(backward), first_slot, num_slots, _stream()), "cx_rope_inplace")
    return qkv


def swiglu_fwd(yg):
    T, two_i = yg.shape
    out = torch.empty(T, two_i // 2, device=yg.device, dtype=torch.bfloat16)
    lib = _lib.load()
    _lib.check(lib.cx_swiglu_fwd(yg.data_ptr(), out.data_ptr(), T, two_i // 2, _stream()), "cx_swiglu_fwd")
    return out


def swiglu_bwd(dout, yg):
    T, two_i = yg.shape
    dyg = torch.empty_like(yg)
    lib = _lib.load()
    _lib.check(lib.cx_swiglu_bwd(dout.data_ptr(), yg.data_ptr(), dyg.data_ptr(), T, two_i // 2, _stream()), "cx_swiglu_bwd")
    return dyg


def mean_pool_fwd(h, cu_seqlens):
    nseq = cu_seqlens.numel() - 1
    d = h.shape[1]
    pooled = torch.empty(nseq, d, device=h.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.cx_mean_pool_fwd(h.data_ptr(), cu_seqlens.data_ptr(), pooled.data_ptr(), nseq, d, _stream()),
               "cx_mean_pool_fwd")
    return pooled


def mean_pool_bwd(dpooled, cu_seqlens, total_tokens):
    nseq, d = dpooled.shape
    dh = torch.empty(total_tokens, d, device=dpooled.device, dtype=torch.bfloat16)
    lib = _lib.load()
    _lib.check(lib.cx_mean_pool_bwd(dpooled.data_ptr(), cu_seqlens.data_ptr(), dh.data_ptr(), nseq, d, _stream()),
               "cx_mean_pool_bwd")
    return dh


def embed_head_fwd(pooled, hamming, normalize):
    rows, d = pooled.shape
    out = torch.empty_like(pooled)
    save = torch.empty(rows, 3, device=pooled.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.cx_embed_head_fwd(pooled.data_ptr(), out.data_ptr(), save.data_ptr(), rows, d, int(hamming),
                                     int(normalize), _stream()), "cx_embed_head_fwd")
    return out, save


def embed_head_bwd(pooled, gout, save, hamming, normalize):
    rows, d = pooled.shape
    g = torch.empty_like(pooled)
    lib = _lib.load()
    _lib.check(lib.cx_embed_head_bwd(pooled.data_ptr(), gout.data_ptr(), save.data_ptr(), g.data_ptr(), rows, d,
                                     int(hamming), int(normalize), _stream()), "cx_embed_head_bwd")
    return g


def attn_fwd(qkv, cu_seqlens, max_seqlen, H, Dh, softmax_scale):
    """qkv [T, 3*H*Dh] bf16 (RoPE applied) -> (out [T, H*Dh] bf16, lse [H, T] fp32)."""
    T = qkv.shape[0]
    nseq = cu_seqlens.numel() - 1
    out = torch.empty(T, H * Dh, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(H, T, device=qkv.device, dtype=torch.float32)
    lib = _lib.load()
    ev = TIMER.begin("attn_fwd") if TIMER is not None else None
    _lib.check(lib.cx_attn_fwd(qkv.data_ptr(), cu_seqlens.data_ptr(), out.data_ptr(), lse.data_ptr(), T, nseq,
                               int(max_seqlen), H, Dh, float(softmax_scale), _stream()), "cx_attn_fwd")
    if ev is not None:  # dense estimate 4*S*d per token; exact for full-length sequences
        TIMER.end("attn_fwd", 4.0 * T * max_seqlen * H * Dh, ev)
    return out, lse


def attn_bwd(qkv, out, dout, lse, cu_seqlens, max_seqlen, H, Dh, softmax_scale, pos=None, cos_t=None, sin_t=None, inv_freq=None):
    """Returns dqkv [T, 3*H*Dh] bf16; with pos/cos/sin/inv_freq the RoPE transpose is applied to dq (in the pass that converts
    the fp32 dQ accumulator) and to dk (in the attention kernel's epilogue)."""
    T = qkv.shape[0]
    nseq = cu_seqlens.numel() - 1
    dqkv = torch.empty_like(qkv)
    dq_acc = torch.empty(T, H * Dh, device=qkv.device, dtype=torch.float32)  # zeroed by the delta kernel inside cx_attn_bwd
    delta = torch.empty(H, T, device=qkv.device, dtype=torch.float32)
    lib = _lib.load()
    ev = TIMER.begin("attn_bwd") if TIMER is not None else None
    _lib.check(lib.cx_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), cu_seqlens.data_ptr(),
                               dqkv.data_ptr(), dq_acc.data_ptr(), delta.data_ptr(), T, nseq, int(max_seqlen), H, Dh,
                               float(softmax_scale), _ptr(inv_freq) if pos is not None else 0, _stream()), "cx_attn_bwd")
    if pos is not None:
        _lib.check(lib.cx_dq_finalize_rope(dq_acc.data_ptr(), dqkv.data_ptr(), pos.data_ptr(), cos_t.data_ptr(),
                                           sin_t.data_ptr(), T, H, Dh, _stream()), "cx_dq_finalize_rope")
        if inv_freq is None:  # caller without the frequency table: separate rotary pass over the dk slot
            rope_inplace(dqkv, pos, cos_t, sin_t, H, Dh, backward=True, first_slot=1, num_slots=1)
    else:
        _lib.check(lib.cx_dq_finalize(dq_acc.data_ptr(), dqkv.data_ptr(), T, H, Dh, _stream()), "cx_dq_finalize")
    if ev is not None:
        TIMER.end("attn_bwd", 10.0 * T * max_seqlen * H * Dh, ev)
    return dqkv


def grad_clip_coef(grad_flat, max_norm):
    lib = _lib.load()
    out2 = torch.empty(2, device=grad_flat.device, dtype=torch.float32)
    ws = _ws_bytes(lib.cx_grad_clip_workspace_bytes(), grad_flat.device)
    _lib.check(lib.cx_grad_clip_coef(grad_flat.data_ptr(), grad_flat.numel(), float(max_norm), out2.data_ptr(),
                                     ws.data_ptr(), _stream()), "cx_grad_clip_coef")
    return out2


def adamw_step(param, grad, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale_dev=None,
               grad_scale=1.0, zero_grad=True):
    lib = _lib.load()
    _lib.check(lib.cx_adamw_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), _ptr(shadow),
                                 param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                 int(step), _ptr(grad_scale_dev), float(grad_scale), int(zero_grad), _stream()),
               "cx_adamw_step")


def cast_f32_bf16(x, out=None):
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    _lib.check(lib.cx_cast_f32_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "cx_cast_f32_bf16")
    return out


# ------------------------------------------------------------------------------------------------ ViT ops
ACT_GELU, ACT_QUICK_GELU = 0, 1


def linear_bias(x, w, bias):
    """y = x w^T + bias (bias fp32 or None) on the tcgen05 GEMM."""
    if bias is None:
        return gemm(x, w)
    _require_cuda(x, w, bias)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    ev = TIMER.begin("gemm") if TIMER is not None else None
    _lib.check(lib.cx_linear_bias_bf16(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), M, N, K, x.stride(0),
                                       w.stride(0), y.stride(0), _stream()), "cx_linear_bias_bf16")
    if ev is not None:
        TIMER.end("gemm", 2.0 * M * N * K, ev)
    return y


def colsum_into(x, out):
    lib = _lib.load()
    _lib.check(lib.cx_colsum_bf16(x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), _stream()), "cx_colsum_bf16")


def act_fwd(x, kind):
    y = torch.empty_like(x)
    _lib.check(_lib.load().cx_act_fwd(x.data_ptr(), y.data_ptr(), x.numel(), kind, _stream()), "cx_act_fwd")
    return y


def act_bwd(dy, x, kind):
    dx = torch.empty_like(x)
    _lib.check(_lib.load().cx_act_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), kind, _stream()), "cx_act_bwd")
    return dx


def patchify(pixels, patch):
    B, C, H, W = pixels.shape
    px = pixels.float().contiguous()
    K = C * patch * patch
    rows = B * (H // patch) * (W // patch)
    if K % 8 == 0:
        out = torch.empty(rows, K, device=px.device, dtype=torch.bfloat16)
        _lib.check(_lib.load().cx_patchify(px.data_ptr(), out.data_ptr(), B, C, H, W, patch, _stream()), "cx_patchify")
        return out
    tmp = torch.empty(rows, K, device=px.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cx_patchify(px.data_ptr(), tmp.data_ptr(), B, C, H, W, patch, _stream()), "cx_patchify")
    out = torch.zeros(rows, (K + 7) // 8 * 8, device=px.device, dtype=torch.bfloat16)
    out[:, :K] = tmp
    return out[:, :K]


def vit_assemble_fwd(proj, cls, pos, B, nP):
    d = proj.shape[1]
    z = torch.empty(B * (nP + 1), d, device=proj.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cx_vit_assemble_fwd(proj.data_ptr(), cls.data_ptr(), pos.data_ptr(), z.data_ptr(), B, nP, d, _stream()),
               "cx_vit_assemble_fwd")
    return z


def vit_assemble_bwd(dz, dcls, dpos, B, nP):
    d = dz.shape[1]
    dproj = torch.empty(B * nP, d, device=dz.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cx_vit_assemble_bwd(dz.data_ptr(), dproj.data_ptr(), dcls.data_ptr(), dpos.data_ptr(), B, nP, d,
                                               _stream()), "cx_vit_assemble_bwd")
    return dproj


def cls_select_fwd(h, B, S):
    d = h.shape[1]
    out = torch.empty(B, d, device=h.device, dtype=torch.float32)
    _lib.check(_lib.load().cx_cls_select_fwd(h.data_ptr(), out.data_ptr(), B, S, d, _stream()), "cx_cls_select_fwd")
    return out


def cls_select_bwd(g, B, S):
    d = g.shape[1]
    dh = torch.empty(B * S, d, device=g.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cx_cls_select_bwd(g.data_ptr(), dh.data_ptr(), B, S, d, _stream()), "cx_cls_select_bwd")
    return dh
