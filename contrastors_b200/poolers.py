"""Poolers and the projection head of the reference's BiEncoder (models/biencoder/modeling_biencoder.py:44-152, 270-285, 312).

* ``MeanPooling`` / ``ClsSelector`` live inside the trunk's autograd node (``_TrunkFn`` / ``_ViTFn``: cx_mean_pool_*, row gather).
* ``MultiHeadAttentionPooling`` ("map", the vision tower of configs/train/nomic_embed_vision_v1.5.yaml:69): one learned latent
  query attends over each image's tokens (FlashAttentionPooling, layers/attention.py:313-440), then
  ``hidden[:, 0] + mlp(norm1(attn))`` (:140-152).  The token-sized work runs on the sm_100a kernels: the Wkv projection is the
  tcgen05 GEMM, the single-query attention is ``cx_attn_pool_fwd/bwd`` (kv read once); what is left operates on B rows
  (out_proj, LayerNorm, the MLP on [B, d]) and goes through the same GEMM with small-M tiles and a few row-sized torch ops.
* ``proj`` (``nn.Linear(hidden, projection_dim)``, :270-273, applied after pooling / hamming and before normalize, :312).

State-dict keys follow the reference: ``selector.attn.Wq|Wkv|out_proj.weight|bias``, ``selector.attn.latent``,
``selector.mlp.fc1|fc2.*`` (``fc11|fc12|fc2`` for GLU activations), ``selector.norm1.*``, ``proj.weight|bias``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .ops import MAJOR_MN


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with x bf16 [M, K], W fp32 master [N, K] (cast to bf16 per call: these are small head matrices), b fp32 or
    None; forward and both gradients on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        wb = weight.detach().to(torch.bfloat16).contiguous()
        x = x.contiguous()
        y = ops.linear_bias(x, wb, bias.detach().float().contiguous() if bias is not None else None)
        ctx.save_for_backward(x, wb)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        dy = dy.contiguous().to(torch.bfloat16)
        dx = ops.gemm(dy, wb, b_major=MAJOR_MN) if ctx.needs_input_grad[0] else None
        dw = ops.gemm(dy, x, a_major=MAJOR_MN, b_major=MAJOR_MN, out_dtype=torch.float32)
        db = dy.float().sum(0) if ctx.has_bias else None
        return dx, dw, db


def linear(x_bf16, module: nn.Linear):
    return _LinearFn.apply(x_bf16, module.weight, module.bias)


class _AttnPoolFn(torch.autograd.Function):
    """out[b, h] = softmax(q_h . K_b^T / sqrt(Dh)) V_b for ONE query per head shared by all sequences (cx_attn_pool_*)."""

    @staticmethod
    def forward(ctx, q, kv, cu, max_seqlen, H):
        nseq = cu.numel() - 1
        Dh = q.numel() // H
        scale = 1.0 / math.sqrt(Dh)
        q = q.detach().float().contiguous()
        kv = kv.contiguous()
        out = torch.empty(nseq, H * Dh, device=kv.device, dtype=torch.float32)
        lse = torch.empty(nseq, H, device=kv.device, dtype=torch.float32)
        lib = _lib.load()
        _lib.check(lib.cx_attn_pool_fwd(q.data_ptr(), kv.data_ptr(), cu.data_ptr(), out.data_ptr(), lse.data_ptr(), nseq,
                                        int(max_seqlen), H, Dh, scale, ops._stream()), "cx_attn_pool_fwd")
        ctx.save_for_backward(q, kv, cu, lse)
        ctx.meta = (nseq, int(max_seqlen), H, Dh, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, cu, lse = ctx.saved_tensors
        nseq, max_seqlen, H, Dh, scale = ctx.meta
        dq = torch.zeros_like(q)
        dkv = torch.empty_like(kv)
        lib = _lib.load()
        _lib.check(lib.cx_attn_pool_bwd(q.data_ptr(), kv.data_ptr(), cu.data_ptr(), dout.contiguous().float().data_ptr(),
                                        lse.data_ptr(), dq.data_ptr(), dkv.data_ptr(), nseq, max_seqlen, H, Dh, scale,
                                        ops._stream()), "cx_attn_pool_bwd")
        return dq, dkv, None, None, None


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


class _PoolAttention(nn.Module):
    """Parameters of FlashAttentionPooling under the reference's names (layers/attention.py:324-354)."""

    def __init__(self, d, bias=True):
        super().__init__()
        self.Wq = nn.Linear(d, d, bias=bias)
        self.Wkv = nn.Linear(d, 2 * d, bias=bias)
        self.latent = nn.Parameter(torch.zeros(1, 1, d))
        self.out_proj = nn.Linear(d, d, bias=bias)
        nn.init.trunc_normal_(self.latent, std=d ** -0.5)


class _PoolMLP(nn.Module):
    def __init__(self, d, inner, gated, bias1=True, bias2=True):
        super().__init__()
        if gated:
            self.fc11 = nn.Linear(d, inner, bias=bias1)
            self.fc12 = nn.Linear(d, inner, bias=bias1)
        else:
            self.fc1 = nn.Linear(d, inner, bias=bias1)
        self.fc2 = nn.Linear(inner, d, bias=bias2)


class MultiHeadAttentionPooling(nn.Module):
    """``selector`` for pooling='map' (modeling_biencoder.py:93-152) on dense token grids (the vision towers: every sequence at
    full length, attention_mask None).  forward(hidden [B*S, d] bf16, B, S) -> [B, d] fp32."""

    def __init__(self, n_embd, n_head, n_inner, activation_function="gelu", layer_norm_epsilon=1e-5, qkv_proj_bias=True,
                 mlp_fc1_bias=True, mlp_fc2_bias=True):
        super().__init__()
        assert n_embd // n_head == 64, "cx_attn_pool_* is specialised for head_dim 64"
        self.n_embd, self.n_head, self.act = n_embd, n_head, activation_function
        self.gated = activation_function in ("glu", "swiglu")
        self.attn = _PoolAttention(n_embd, bias=qkv_proj_bias)
        self.mlp = _PoolMLP(n_embd, n_inner, self.gated, mlp_fc1_bias, mlp_fc2_bias)
        self.norm1 = nn.LayerNorm(n_embd, eps=layer_norm_epsilon)

    def _activation(self, x):
        return {"glu": torch.sigmoid, "swiglu": F.silu, "quick_gelu": quick_gelu}.get(self.act, F.gelu)(x)

    def forward(self, hidden, B, S):
        d, H = self.n_embd, self.n_head
        cu = torch.arange(0, (B + 1) * S, S, device=hidden.device, dtype=torch.int32)
        q = F.linear(self.attn.latent.reshape(1, d).float(), self.attn.Wq.weight.float(),
                     None if self.attn.Wq.bias is None else self.attn.Wq.bias.float()).reshape(-1)          # [d], one row: torch
        kv = linear(hidden, self.attn.Wkv)                                                                  # [T, 2d] tcgen05
        pooled = _AttnPoolFn.apply(q, kv, cu, S, H)                                                         # [B, d] fp32
        o = linear(pooled.to(torch.bfloat16), self.attn.out_proj).float()
        normed = F.layer_norm(o, (d,), self.norm1.weight.float(), self.norm1.bias.float(), self.norm1.eps).to(torch.bfloat16)
        if self.gated:
            y = linear(normed, self.mlp.fc11).float() * self._activation(linear(normed, self.mlp.fc12).float())
        else:
            y = self._activation(linear(normed, self.mlp.fc1).float())
        m = linear(y.to(torch.bfloat16), self.mlp.fc2).float()
        return hidden.view(B, S, d)[:, 0].float() + m   # hidden_states + mlp(normed) broadcast, then [:, 0] (:148-152)


def head_parameters(model):
    """Trainable parameters of a tower that live outside the trunk's flat buffers (selector.*, proj.*)."""
    return [p for n, p in model.named_parameters() if not n.startswith("trunk.") and p.requires_grad]
