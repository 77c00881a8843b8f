"""ViT oracle: explicit fp32 torch-CPU restatement of the reference's CLIP-style vision tower.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/contrastors/models/vit/vit.py:176-276 (ViTModel.forward), layers/embedding.py:465-516
(PatchEmbedding: patch rearrange "(c p1 p2)" -> Linear, cls token, learned position embedding), layers/block.py:293-388
(pre-norm Block: x += attn(LN1(x)); x += mlp(LN2(x))), layers/mlp.py:8-34 (MLP with biases, gelu / quick_gelu) and the
CLIP config mapping models/vit/clip.py:9-53 (prepre LayerNorm, qkv/mlp biases, no patch bias, ln_f, CLS pooling).
State-dict keys are the reference's (models/vit/clip.py:56-173).  Pinned against transformers.CLIPVisionModel through the
reference's own remap (oracle/gen_golden.py::gen_vit), exactly how the reference tests it (tests/test_flash_openclip.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class ViTConfig:
    n_embd: int = 768
    n_head: int = 12
    n_inner: int = 3072
    n_layer: int = 12
    img_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    activation_function: str = "quick_gelu"
    layer_norm_epsilon: float = 1e-5

    @property
    def num_patches(self):
        return (self.img_size // self.patch_size) ** 2


def state_dict_keys(cfg: ViTConfig):
    d, I = cfg.n_embd, cfg.n_inner
    keys = [("embeddings.proj.weight", (d, cfg.num_channels * cfg.patch_size ** 2)), ("embeddings.cls_token", (1, 1, d)),
            ("embeddings.pos_embed", (1, cfg.num_patches + 1, d)), ("prepre_layernom.weight", (d,)), ("prepre_layernom.bias", (d,))]
    for i in range(cfg.n_layer):
        p = f"layers.{i}."
        keys += [(p + "norm1.weight", (d,)), (p + "norm1.bias", (d,)), (p + "attn.Wqkv.weight", (3 * d, d)),
                 (p + "attn.Wqkv.bias", (3 * d,)), (p + "attn.out_proj.weight", (d, d)), (p + "attn.out_proj.bias", (d,)),
                 (p + "norm2.weight", (d,)), (p + "norm2.bias", (d,)), (p + "mlp.fc1.weight", (I, d)), (p + "mlp.fc1.bias", (I,)),
                 (p + "mlp.fc2.weight", (d, I)), (p + "mlp.fc2.bias", (d,))]
    keys += [("ln_f.weight", (d,)), ("ln_f.bias", (d,))]
    return keys


def random_state_dict(cfg: ViTConfig, seed=0):
    rs = np.random.RandomState(seed)
    sd = {}
    for k, shape in state_dict_keys(cfg):
        if "norm" in k or k.startswith("ln_f") or k.startswith("prepre"):
            v = (1.0 + 0.1 * rs.randn(*shape)) if k.endswith("weight") else 0.1 * rs.randn(*shape)
        elif k.endswith(".bias"):
            v = 0.05 * rs.randn(*shape)
        else:
            v = 0.03 * rs.randn(*shape)
        sd[k] = torch.from_numpy(v.astype(np.float32))
    return sd


def _act(x, kind):
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    return F.gelu(x)


def vit_forward(sd, cfg: ViTConfig, pixels, dtype=torch.float32):
    """pixels [B, C, H, W] -> CLS embedding after ln_f [B, d] (BiEncoder pooling='cls', normalize=False)."""
    g = lambda k: sd[k].to(dtype)
    B, C, H, W = pixels.shape
    p = cfg.patch_size
    x = pixels.to(dtype).reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
    x = x @ g("embeddings.proj.weight").T
    x = torch.cat([g("embeddings.cls_token").reshape(1, 1, -1).expand(B, 1, -1), x], dim=1) + g("embeddings.pos_embed")
    x = F.layer_norm(x, (cfg.n_embd,), g("prepre_layernom.weight"), g("prepre_layernom.bias"), cfg.layer_norm_epsilon)
    Hh, Dh = cfg.n_head, cfg.n_embd // cfg.n_head
    S = x.shape[1]
    for i in range(cfg.n_layer):
        q = f"layers.{i}."
        h = F.layer_norm(x, (cfg.n_embd,), g(q + "norm1.weight"), g(q + "norm1.bias"), cfg.layer_norm_epsilon)
        qkv = (h @ g(q + "attn.Wqkv.weight").T + g(q + "attn.Wqkv.bias")).view(B, S, 3, Hh, Dh)
        qq, kk, vv = (qkv[:, :, j].permute(0, 2, 1, 3) for j in range(3))
        a = torch.softmax((qq @ kk.transpose(-1, -2)) / math.sqrt(Dh), dim=-1) @ vv
        a = a.permute(0, 2, 1, 3).reshape(B, S, cfg.n_embd) @ g(q + "attn.out_proj.weight").T + g(q + "attn.out_proj.bias")
        x = x + a
        h = F.layer_norm(x, (cfg.n_embd,), g(q + "norm2.weight"), g(q + "norm2.bias"), cfg.layer_norm_epsilon)
        m = _act(h @ g(q + "mlp.fc1.weight").T + g(q + "mlp.fc1.bias"), cfg.activation_function)
        x = x + (m @ g(q + "mlp.fc2.weight").T + g(q + "mlp.fc2.bias"))
    x = F.layer_norm(x, (cfg.n_embd,), g("ln_f.weight"), g("ln_f.bias"), cfg.layer_norm_epsilon)
    return x[:, 0]
