"""Encoder oracle: explicit fp32/fp64 torch-CPU restatement of the reference's nomic-bert tower.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's own pure-PyTorch model (the reference tests assert it has identical
state-dict keys/shapes to the flash model, tests/test_huggingface.py:30-34):
  * embeddings        models/huggingface/modeling_hf_nomic_bert.py:962-1000  (word + token-type, no abs-pos with rotary)
  * emb_ln / dropout  :1650-1695
  * rotary (NeoX)     :1074-1212   cos/sin computed in fp32 then cast to the activation dtype
  * attention         :1285-1414   softmax(QK^T/sqrt(Dh) + mask) V, non-causal
  * gated MLP         :1031-1071   fc2( fc11(x) * silu(fc12(x)) )
  * post-norm block   :1497-1514   h = LN1(attn(h)+h); h = LN2(mlp(h)+h)
  * BiEncoder tail    models/biencoder/modeling_biencoder.py:79-90 (mean pool), :282-285,307 (hamming LN),
                      :309-317 (cast back to trunk dtype, F.normalize)
State-dict keys are the reference's ("embeddings.word_embeddings.weight", "emb_ln.weight",
"encoder.layers.{i}.attn.Wqkv.weight", "...attn.out_proj.weight", "...mlp.fc11.weight", "...mlp.fc12.weight",
"...mlp.fc2.weight", "...norm1.weight", ...).  Dropout is not modelled (p = 0); backward uses torch autograd
on this explicit graph.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class EncoderConfig:
    vocab_size: int = 30528
    n_embd: int = 768
    n_head: int = 12
    n_inner: int = 3072
    n_layer: int = 12
    type_vocab_size: int = 2
    rotary_emb_base: float = 1000.0
    layer_norm_epsilon: float = 1e-12
    initializer_range: float = 0.02

    @property
    def head_dim(self):
        return self.n_embd // self.n_head


def state_dict_keys(cfg: EncoderConfig):
    keys = [("embeddings.word_embeddings.weight", (cfg.vocab_size, cfg.n_embd)),
            ("embeddings.token_type_embeddings.weight", (cfg.type_vocab_size, cfg.n_embd)),
            ("emb_ln.weight", (cfg.n_embd,)), ("emb_ln.bias", (cfg.n_embd,))]
    for i in range(cfg.n_layer):
        p = f"encoder.layers.{i}."
        keys += [(p + "attn.Wqkv.weight", (3 * cfg.n_embd, cfg.n_embd)),
                 (p + "attn.out_proj.weight", (cfg.n_embd, cfg.n_embd)),
                 (p + "mlp.fc11.weight", (cfg.n_inner, cfg.n_embd)),
                 (p + "mlp.fc12.weight", (cfg.n_inner, cfg.n_embd)),
                 (p + "mlp.fc2.weight", (cfg.n_embd, cfg.n_inner)),
                 (p + "norm1.weight", (cfg.n_embd,)), (p + "norm1.bias", (cfg.n_embd,)),
                 (p + "norm2.weight", (cfg.n_embd,)), (p + "norm2.bias", (cfg.n_embd,))]
    return keys


def random_state_dict(cfg: EncoderConfig, seed: int = 0, ln_jitter: float = 0.1):
    """Deterministic weights from the frozen numpy RandomState stream (stable across numpy versions).

    Linear/embedding ~ N(0, initializer_range) as modeling_nomic_bert.py:284-292; LayerNorm gains/biases are
    jittered around (1, 0) so parity tests exercise them.
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for k, shape in state_dict_keys(cfg):
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "emb_ln.weight":
            v = 1.0 + ln_jitter * rs.randn(*shape)
        elif k.endswith(".bias"):
            v = ln_jitter * rs.randn(*shape)
        else:
            v = cfg.initializer_range * rs.randn(*shape)
        sd[k] = torch.from_numpy(v.astype(np.float32))
    return sd


def rotary_cos_sin(seqlen: int, dim: int, base: float, dtype=torch.float32):
    """modeling_hf_nomic_bert.py:1148-1183: inv_freq fp32, outer product fp32, cos/sin cast to dtype."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    t = torch.arange(seqlen, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype)


def apply_rotary_neox(x, cos, sin):
    """x [B,S,H,Dh]; NeoX/non-interleaved: out = x*cos + rotate_half(x)*sin (:1074-1100)."""
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half], x[..., half:]
    c = cos[None, : x.shape[1], None, :]
    s = sin[None, : x.shape[1], None, :]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if w is not None:
        y = y * w + b
    return y


def trunk_forward(sd, cfg: EncoderConfig, input_ids, attention_mask=None, token_type_ids=None, dtype=torch.float32):
    """NomicBertModel.forward -> last_hidden_state [B,S,d] (modeling_hf_nomic_bert.py:1676-1695)."""
    B, S = input_ids.shape
    g = lambda k: sd[k].to(dtype)
    h = g("embeddings.word_embeddings.weight")[input_ids]
    tt = torch.zeros_like(input_ids) if token_type_ids is None else token_type_ids
    h = h + g("embeddings.token_type_embeddings.weight")[tt]
    h = layer_norm(h, g("emb_ln.weight"), g("emb_ln.bias"), cfg.layer_norm_epsilon)
    dev = input_ids.device  # CPU for the oracle proper; the "plain bf16 run of the same graph" arm of the tests may sit on the GPU
    if attention_mask is None:
        attention_mask = torch.ones(B, S, dtype=torch.long, device=dev)
    bias = torch.zeros(B, 1, 1, S, dtype=dtype, device=dev).masked_fill(attention_mask[:, None, None, :] == 0, float("-inf"))
    cos, sin = (t.to(dev) for t in rotary_cos_sin(S, cfg.head_dim, cfg.rotary_emb_base, dtype))
    H, Dh = cfg.n_head, cfg.head_dim
    for i in range(cfg.n_layer):
        p = f"encoder.layers.{i}."
        qkv = (h @ g(p + "attn.Wqkv.weight").T).view(B, S, 3, H, Dh)
        q = apply_rotary_neox(qkv[:, :, 0], cos, sin).permute(0, 2, 1, 3)
        k = apply_rotary_neox(qkv[:, :, 1], cos, sin).permute(0, 2, 1, 3)
        v = qkv[:, :, 2].permute(0, 2, 1, 3)
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(Dh) + bias
        a = torch.softmax(scores, dim=-1) @ v
        a = a.permute(0, 2, 1, 3).reshape(B, S, H * Dh) @ g(p + "attn.out_proj.weight").T
        h = layer_norm(a + h, g(p + "norm1.weight"), g(p + "norm1.bias"), cfg.layer_norm_epsilon)
        y = (h @ g(p + "mlp.fc11.weight").T) * F.silu(h @ g(p + "mlp.fc12.weight").T)
        m = y @ g(p + "mlp.fc2.weight").T
        h = layer_norm(m + h, g(p + "norm2.weight"), g(p + "norm2.bias"), cfg.layer_norm_epsilon)
    return h


def biencoder_forward(sd, cfg: EncoderConfig, input_ids, attention_mask=None, normalize=True, hamming=False,
                      dtype=torch.float32):
    """BiEncoder.forward with pooling='mean' (modeling_biencoder.py:287-319) -> embedding [B,d]."""
    h = trunk_forward(sd, cfg, input_ids, attention_mask, dtype=dtype)
    if attention_mask is None:
        e = h.mean(dim=1)
    else:
        m = attention_mask.unsqueeze(-1).float()
        e = (h * m).sum(dim=1) / attention_mask.sum(dim=1, keepdim=True).float()
    if hamming:
        e = layer_norm(e, None, None, 1e-5)
    e = e.to(h.dtype)
    if normalize:
        e = F.normalize(e, dim=-1)
    return e
