"""CLIP-style ViT tower and the two-tower DualEncoder on the sm_100a kernels (configs 3 / 5 of BASELINE.json).

Reference: models/vit/vit.py:107-276 (ViTModel), layers/embedding.py:357-516 (PatchEmbedding), layers/block.py:293-388
(pre-norm Block), layers/mlp.py:8-34 (MLP + FusedDense biases), models/vit/clip.py:9-53 (CLIP -> ViT config: prepre LN,
qkv / mlp biases, no patch bias, ln_f), models/biencoder/modeling_biencoder.py:44-49 (ClsSelector),
models/dual_encoder/modeling_dual_encoder.py:10-68 (DualEncoder).  Parameter names are the reference's
(models/vit/clip.py:56-173): ``embeddings.proj.weight``, ``embeddings.cls_token``, ``embeddings.pos_embed``,
``prepre_layernom.*``, ``layers.{i}.norm1|attn.Wqkv|attn.out_proj|norm2|mlp.fc1|mlp.fc2.*``, ``ln_f.*``.

Same storage design as the text tower: one flat fp32 master / gradient / bf16 shadow, one autograd node per tower call.
The residual stream is carried in bf16 between the fused add-LayerNorm kernels (z_out), attention is the same varlen
tcgen05 kernel with every sequence at full length (197 / 257 tokens -> a 128-row tile plus a masked remainder).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .loss import symmetric_clip_loss
from .logit_scale import LogitScale
from .flat_params import FlatParamModule
from .ops import MAJOR_MN


@dataclass
class ViTConfig:
    n_embd: int = 768
    n_head: int = 12
    n_inner: int = 3072
    n_layer: int = 12
    img_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    activation_function: str = "quick_gelu"   # OpenAI CLIP; "gelu" for the others
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02

    @property
    def num_patches(self):
        return (self.img_size // self.patch_size) ** 2

    @property
    def hidden_size(self):
        return self.n_embd


def vit_b16(**kw):
    return ViTConfig(**kw)


def vit_l14(**kw):
    return ViTConfig(n_embd=1024, n_head=16, n_inner=4096, n_layer=24, patch_size=14, **kw)


def _vit_specs(cfg: ViTConfig):
    d, I = cfg.n_embd, cfg.n_inner
    two_d = [("embeddings.proj.weight", (d, cfg.num_channels * cfg.patch_size ** 2)), ("embeddings.pos_embed", (1, cfg.num_patches + 1, d))]
    one_d = [("embeddings.cls_token", (1, 1, d)), ("prepre_layernom.weight", (d,)), ("prepre_layernom.bias", (d,))]
    for i in range(cfg.n_layer):
        p = f"layers.{i}."
        two_d += [(p + "attn.Wqkv.weight", (3 * d, d)), (p + "attn.out_proj.weight", (d, d)), (p + "mlp.fc1.weight", (I, d)),
                  (p + "mlp.fc2.weight", (d, I))]
        one_d += [(p + "norm1.weight", (d,)), (p + "norm1.bias", (d,)), (p + "attn.Wqkv.bias", (3 * d,)),
                  (p + "attn.out_proj.bias", (d,)), (p + "norm2.weight", (d,)), (p + "norm2.bias", (d,)),
                  (p + "mlp.fc1.bias", (I,)), (p + "mlp.fc2.bias", (d,))]
    one_d += [("ln_f.weight", (d,)), ("ln_f.bias", (d,))]
    return two_d, one_d


class ViTModel(FlatParamModule):
    def __init__(self, config: ViTConfig):
        super().__init__()
        self.config = config
        assert config.n_embd // config.n_head == 64, "the sm_100a attention kernel is specialised for head_dim 64"
        two_d, one_d = _vit_specs(config)
        self._init_flat(two_d, one_d)
        self.reset_parameters()

    def reset_parameters(self, seed: Optional[int] = None):
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            for name, (off, n, shape) in self._offsets.items():
                v = self._flat[off:off + n]
                if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name in ("ln_f.weight", "prepre_layernom.weight"):
                    v.fill_(1.0)
                elif name.endswith(".bias") or name == "embeddings.cls_token":
                    v.zero_()
                else:
                    v.copy_((torch.randn(n, generator=g) * self.config.initializer_range).to(v.device))
        self.mark_weights_updated()

    def forward(self, input_ids, **kwargs):
        """``input_ids`` = pixel tensor [B, C, H, W] (the reference keys pixels as input_ids, image_text_loader.py:339).
        Returns (last_hidden_state [B, S, d],) after ln_f."""
        B = input_ids.shape[0]
        h = _ViTFn.apply(self._flat, self, input_ids, False)
        return (h.view(B, self.config.num_patches + 1, -1),)


class _ViTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, model: ViTModel, pixels, pooled: bool):
        cfg = model.config
        need_grad = ctx.needs_input_grad[0]
        W, P, v = model.shadow(), model._flat, model.view
        d, H, Dh, nP = cfg.n_embd, cfg.n_head, 64, cfg.num_patches
        S = nP + 1
        B = pixels.shape[0]
        T = B * S
        eps = cfg.layer_norm_epsilon
        act = ops.ACT_QUICK_GELU if cfg.activation_function == "quick_gelu" else ops.ACT_GELU
        scale = 1.0 / math.sqrt(Dh)
        cu = torch.arange(0, (B + 1) * S, S, device=pixels.device, dtype=torch.int32)
        patches = ops.patchify(pixels, cfg.patch_size)   # [B*nP, C*p*p] (row stride padded to 8 elements for TMA)
        wproj = v(W, "embeddings.proj.weight")
        if wproj.shape[1] % 8 != 0:  # e.g. patch 14: 588 columns -> 16-byte aligned padded copy of the (small) weight
            kp = patches.stride(0)
            wpad = torch.zeros(wproj.shape[0], kp, device=wproj.device, dtype=torch.bfloat16)
            wpad[:, :wproj.shape[1]] = wproj
            wproj = wpad[:, :wproj.shape[1]]
        proj = ops.gemm(patches, wproj)
        z0 = ops.vit_assemble_fwd(proj, v(P, "embeddings.cls_token").reshape(-1), v(P, "embeddings.pos_embed").reshape(S, d), B, nP)
        x, st_pre = ops.add_layernorm_fwd(z0, None, v(P, "prepre_layernom.weight"), v(P, "prepre_layernom.bias"), eps)
        # layer 0: h = LN1(x); later layers: (x, h) = add_ln(m_prev, x_prev) fused
        saved = []
        a_in, b_in = x, None  # the residual stream entering LN1 of this layer is z = a_in + b_in
        for i in range(cfg.n_layer):
            p = f"layers.{i}."
            if b_in is None:
                h, st1 = ops.add_layernorm_fwd(a_in, None, v(P, p + "norm1.weight"), v(P, p + "norm1.bias"), eps)
                xr = a_in
            else:
                h, st1, xr = ops.add_layernorm_fwd(a_in, b_in, v(P, p + "norm1.weight"), v(P, p + "norm1.bias"), eps, want_z=True)
            qkv = ops.linear_bias(h, v(W, p + "attn.Wqkv.weight"), v(P, p + "attn.Wqkv.bias"))
            attn, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
            o = ops.linear_bias(attn, v(W, p + "attn.out_proj.weight"), v(P, p + "attn.out_proj.bias"))
            h2, st2, x1 = ops.add_layernorm_fwd(o, xr, v(P, p + "norm2.weight"), v(P, p + "norm2.bias"), eps, want_z=True)
            y = ops.linear_bias(h2, v(W, p + "mlp.fc1.weight"), v(P, p + "mlp.fc1.bias"))
            a = ops.act_fwd(y, act)
            m = ops.linear_bias(a, v(W, p + "mlp.fc2.weight"), v(P, p + "mlp.fc2.bias"))
            if need_grad:
                saved.append((a_in, b_in, st1, h, qkv, attn, lse, o, xr, st2, h2, y, a))
            a_in, b_in = m, x1
        out, st_f = ops.add_layernorm_fwd(a_in, b_in, v(P, "ln_f.weight"), v(P, "ln_f.bias"), eps)
        ctx.model, ctx.saved, ctx.pooled = model, saved, pooled
        ctx.misc = (patches, z0, st_pre, a_in, b_in, st_f, cu, B, S, act)
        if pooled:
            return ops.cls_select_fwd(out, B, S)  # ClsSelector: hidden_states[:, 0] (fp32)
        return out

    @staticmethod
    def backward(ctx, g_out):
        model = ctx.model
        cfg = model.config
        model._ensure_grad_views()
        W, P, G, v = model.shadow(), model._flat, model._flat_grad, model.view
        patches, z0, st_pre, a_f, b_f, st_f, cu, B, S, act = ctx.misc
        d, H, Dh, nP = cfg.n_embd, cfg.n_head, 64, cfg.num_patches
        scale = 1.0 / math.sqrt(Dh)
        g = ops.cls_select_bwd(g_out.contiguous().float(), B, S) if ctx.pooled else g_out.contiguous().to(torch.bfloat16)
        # ln_f: z = m_last + x1_last
        dz = ops.add_layernorm_bwd(a_f, b_f, g, None, v(P, "ln_f.weight"), st_f, v(G, "ln_f.weight"), v(G, "ln_f.bias"))
        for i in reversed(range(cfg.n_layer)):
            p = f"layers.{i}."
            a_in, b_in, st1, h, qkv, attn, lse, o, xr, st2, h2, y, a = ctx.saved[i]
            ctx.saved[i] = None
            # dz = gradient of the residual stream after this layer's MLP add: flows to m (MLP) and to x1
            dm = dz
            ops.colsum_into(dm, v(G, p + "mlp.fc2.bias"))
            da = ops.gemm(dm, v(W, p + "mlp.fc2.weight"), b_major=MAJOR_MN)
            ops.gemm(dm, a, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "mlp.fc2.weight"), accumulate=True)
            dy = ops.act_bwd(da, y, act)
            ops.colsum_into(dy, v(G, p + "mlp.fc1.bias"))
            dh2 = ops.gemm(dy, v(W, p + "mlp.fc1.weight"), b_major=MAJOR_MN)
            ops.gemm(dy, h2, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "mlp.fc1.weight"), accumulate=True)
            # x1 = o + xr feeds LN2 and the residual stream (gres = dz)
            dz1 = ops.add_layernorm_bwd(o, xr, dh2, None, v(P, p + "norm2.weight"), st2, v(G, p + "norm2.weight"),
                                        v(G, p + "norm2.bias"), gres=dz)
            ops.colsum_into(dz1, v(G, p + "attn.out_proj.bias"))
            dattn = ops.gemm(dz1, v(W, p + "attn.out_proj.weight"), b_major=MAJOR_MN)
            ops.gemm(dz1, attn, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "attn.out_proj.weight"), accumulate=True)
            dqkv = ops.attn_bwd(qkv, attn, dattn, lse, cu, S, H, Dh, scale)
            ops.colsum_into(dqkv, v(G, p + "attn.Wqkv.bias"))
            dh = ops.gemm(dqkv, v(W, p + "attn.Wqkv.weight"), b_major=MAJOR_MN)
            ops.gemm(dqkv, h, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "attn.Wqkv.weight"), accumulate=True)
            # xr = a_in (+ b_in) feeds LN1 and the residual stream (gres = dz1)
            dz = ops.add_layernorm_bwd(a_in, b_in, dh, None, v(P, p + "norm1.weight"), st1, v(G, p + "norm1.weight"),
                                       v(G, p + "norm1.bias"), gres=dz1)
        # dz is now the gradient of x = prepre_LN(z0)
        dz0 = ops.add_layernorm_bwd(z0, None, dz, None, v(P, "prepre_layernom.weight"), st_pre, v(G, "prepre_layernom.weight"),
                                    v(G, "prepre_layernom.bias"))
        dproj = ops.vit_assemble_bwd(dz0, v(G, "embeddings.cls_token").reshape(-1), v(G, "embeddings.pos_embed").reshape(S, d), B, nP)
        ops.gemm(dproj, patches, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, "embeddings.proj.weight"), accumulate=True)
        ctx.saved = None
        return None, None, None, None


@dataclass
class VisionBiEncoderConfig:
    """Vision-side BiEncoderConfig fields this path reads (pooling='cls', no projection, optional freeze)."""
    pooling: str = "cls"
    freeze: bool = False
    projection_dim: Optional[int] = None
    logit_scale: float = 1.0 / 0.07
    trainable_logit_scale: bool = True
    encoder: Optional[ViTConfig] = None


class VisionBiEncoder(nn.Module):
    """BiEncoder around the ViT trunk (modeling_biencoder.py:155-319 with a CLIP trunk, ClsSelector, Identity proj)."""

    chunk_streams_ok = False

    def __init__(self, config: VisionBiEncoderConfig):
        super().__init__()
        self.config = config
        if config.pooling not in ("cls", "map"):
            raise NotImplementedError("vision tower: pooling='cls' (CLIP) or 'map' (configs/train/nomic_embed_vision_v1.5.yaml:69)")
        self.trunk = ViTModel(config.encoder or vit_b16())
        enc = self.trunk.config
        if config.pooling == "map":
            from .poolers import MultiHeadAttentionPooling
            self.selector = MultiHeadAttentionPooling(enc.n_embd, enc.n_head, enc.n_inner, enc.activation_function, enc.layer_norm_epsilon)
        else:
            self.selector = None
        self.proj = nn.Linear(enc.n_embd, config.projection_dim) if config.projection_dim else None
        self.frozen_trunk = bool(config.freeze)
        if self.frozen_trunk:
            self.trunk.eval()
            for p in self.trunk.parameters():
                p.requires_grad = False

    def no_sync(self):
        from contextlib import nullcontext
        return nullcontext()

    def forward(self, input_ids, attention_mask=None, is_padded_inputs=True, normalize=True, binarize=False, **kwargs):
        flat = self.trunk._flat
        pooled_in_trunk = self.selector is None  # ClsSelector runs inside the trunk's node; MAP needs every token
        if self.frozen_trunk:
            with torch.no_grad():
                emb = _ViTFn.apply(flat, self.trunk, input_ids, pooled_in_trunk)
        else:
            if not flat.requires_grad:
                flat.requires_grad_(True)
            emb = _ViTFn.apply(flat, self.trunk, input_ids, pooled_in_trunk)
        if self.selector is not None:
            B = input_ids.shape[0]
            emb = self.selector(emb, B, self.trunk.config.num_patches + 1)
        # the reference casts the pooled vector back to the trunk dtype before proj / normalize (modeling_biencoder.py:309-317)
        emb = emb + (emb.to(torch.bfloat16).float() - emb).detach()
        if self.proj is not None:
            from .poolers import linear
            emb = linear(emb.to(torch.bfloat16), self.proj).float()
        if normalize and not binarize:
            emb = torch.nn.functional.normalize(emb, dim=-1)
        if binarize:
            emb = (emb > 0).float()
        return {"embedding": emb, "router_logits": None, "router_loss": None, "tokens_per_expert": None}


class DualEncoder(nn.Module):
    """Two towers + the symmetric CLIP loss (modeling_dual_encoder.py:10-68).  ``forward(text_inputs, vision_inputs)``
    returns {"loss", "image_text_loss"}; ``text_inputs`` may carry precomputed ``text_embs`` for a frozen text tower."""

    def __init__(self, text: nn.Module, vision: nn.Module, logit_scale: float = 1.0 / 0.07, trainable_logit_scale: bool = True,
                 precomputed_text: bool = False):
        super().__init__()
        self.text, self.vision = text, vision
        self.precomputed_text = precomputed_text
        self.logit_scale = LogitScale(logit_scale=logit_scale, trainable_logit_scale=trainable_logit_scale)

    def encode_text(self, text, normalize=True):
        return self.text(**text, normalize=normalize)["embedding"]

    def encode_image(self, vision, normalize=True):
        return self.vision(vision, normalize=normalize)["embedding"]

    def forward(self, text_inputs, vision_inputs):
        if self.precomputed_text:
            assert "text_embs" in text_inputs, "Precomputed text inputs must have text_embs"
            text_emb = text_inputs["text_embs"]
        else:
            text_emb = self.text(**text_inputs, normalize=False)["embedding"]
        vision_emb = self.vision(**vision_inputs, normalize=False)["embedding"]
        loss = symmetric_clip_loss(text_emb, vision_emb, self.logit_scale)
        return {"loss": loss, "image_text_loss": loss}
