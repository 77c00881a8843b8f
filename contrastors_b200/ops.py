"""Tensor-level wrappers over the C ABI (pointer plumbing only; PyTorch is used for device memory and streams)."""
from __future__ import annotations

import torch

from . import _lib

MAJOR_K, MAJOR_MN = 0, 1
_DT = {torch.bfloat16: 0, torch.float32: 1}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("contrastors_b200 ops run on a B200 only: got a CPU tensor (no CPU fallback exists)")


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
    _lib.check(lib.cx_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a_major, b_major, a.stride(0),
                                b.stride(0), out.stride(0), _DT[out.dtype], int(accumulate), float(alpha), _stream()),
               "cx_gemm_bf16")
    return out


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
    lib = _lib.load()
    _lib.check(lib.cx_l2norm_bwd(x.data_ptr(), x.stride(0), g.data_ptr(), g.stride(0), inv_norm.data_ptr(),
                                 out.data_ptr(), out.stride(0), rows, k, int(g_prescaled), int(accumulate), _stream()),
               "cx_l2norm_bwd")
    return out


def infonce_workspace(n, m, device):
    nbytes = _lib.load().cx_infonce_workspace_bytes(n, m)
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
    _lib.check(lib.cx_infonce_fwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, k_dim, float(scale),
                                  _ptr(scale_dev), _ptr(rq), _ptr(rd), label_offset, label_stride, lse.data_ptr(),
                                  argmax.data_ptr(), label_logit.data_ptr(), stats.data_ptr(), workspace.data_ptr(),
                                  _stream()), "cx_infonce_fwd")
    return lse, argmax, label_logit, stats


def infonce_bwd(q, d, k_dim, scale, scale_dev, rq, rd, label_offset, label_stride, lse, coef, coef_dev, dq, dd,
                accumulate_dd, stats, workspace):
    lib = _lib.load()
    n, m = q.shape[0], d.shape[0]
    _lib.check(lib.cx_infonce_bwd(q.data_ptr(), q.stride(0), d.data_ptr(), d.stride(0), n, m, k_dim, float(scale),
                                  _ptr(scale_dev), _ptr(rq), _ptr(rd), label_offset, label_stride, lse.data_ptr(),
                                  float(coef), _ptr(coef_dev), dq.data_ptr(), dq.stride(0), dd.data_ptr(), dd.stride(0),
                                  int(accumulate_dd), stats.data_ptr(), workspace.data_ptr(), _stream()),
               "cx_infonce_bwd")
