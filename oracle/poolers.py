"""Pooler / projection oracle: fp32 torch restatement of the reference's BiEncoder tail (TEST INFRASTRUCTURE ONLY).

  * ``MultiHeadAttentionPooling.forward``  models/biencoder/modeling_biencoder.py:126-152  (attention_mask None: the vision towers)
  * ``FlashAttentionPooling.forward``      layers/attention.py:356-440  (latent query -> Wq; Wkv; softmax(q k^T / sqrt(Dh)) v; out_proj)
  * ``MLP`` / activation choice            modeling_biencoder.py:97-124, layers/mlp.py:8-34
  * ``ClsSelector``                        modeling_biencoder.py:44-49
  * ``proj`` + cast + normalize            modeling_biencoder.py:270-273, 307-317
The flash kernels are replaced by the explicit softmax they compute; parameter names are the reference's.
Pinned: tests/golden/pooler_map_{gelu,swiglu}.npz are outputs and gradients of the reference's own MultiHeadAttentionPooling /
ClsSelector run on CPU (oracle/gen_golden.py::gen_poolers); tests/test_oracle_golden.py holds this file to them.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _act(x, kind):
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if kind == "swiglu":
        return F.silu(x)
    if kind == "glu":
        return torch.sigmoid(x)
    return F.gelu(x)


def map_pool(sd, hidden, n_head, activation="gelu", eps=1e-5):
    """hidden [B, S, d] -> [B, d].  sd keys: attn.latent, attn.Wq.*, attn.Wkv.*, attn.out_proj.*, mlp.fc1.*, mlp.fc2.* (or
    mlp.fc11/fc12/fc2 for gated activations), norm1.*."""
    B, S, d = hidden.shape
    Dh = d // n_head
    g = lambda k: sd[k]
    q = F.linear(g("attn.latent").reshape(1, 1, d).expand(B, 1, d), g("attn.Wq.weight"), sd.get("attn.Wq.bias"))      # [B,1,d]
    kv = F.linear(hidden, g("attn.Wkv.weight"), sd.get("attn.Wkv.bias")).view(B, S, 2, n_head, Dh)
    qh = q.view(B, 1, n_head, Dh).permute(0, 2, 1, 3)
    k, v = kv[:, :, 0].permute(0, 2, 1, 3), kv[:, :, 1].permute(0, 2, 1, 3)
    a = torch.softmax((qh @ k.transpose(-1, -2)) / math.sqrt(Dh), dim=-1) @ v                                          # [B,H,1,Dh]
    a = a.permute(0, 2, 1, 3).reshape(B, 1, d)
    o = F.linear(a, g("attn.out_proj.weight"), sd.get("attn.out_proj.bias"))
    normed = F.layer_norm(o, (d,), g("norm1.weight"), g("norm1.bias"), eps)
    if activation in ("glu", "swiglu"):
        y = F.linear(normed, g("mlp.fc11.weight"), sd.get("mlp.fc11.bias")) * _act(F.linear(normed, g("mlp.fc12.weight"), sd.get("mlp.fc12.bias")), activation)
    else:
        y = _act(F.linear(normed, g("mlp.fc1.weight"), sd.get("mlp.fc1.bias")), activation)
    m = F.linear(y, g("mlp.fc2.weight"), sd.get("mlp.fc2.bias"))
    return (hidden + m)[:, 0]


def project_normalize(pooled, weight, bias, normalize=True, trunk_dtype=torch.bfloat16):
    """embedding.to(trunk dtype) -> proj -> F.normalize (modeling_biencoder.py:309-317)."""
    e = F.linear(pooled.to(trunk_dtype).float(), weight, bias)
    return F.normalize(e, dim=-1) if normalize else e
