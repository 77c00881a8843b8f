"""The reference's own encoder + GradCache loss as a runnable baseline tower (TEST / BENCH INFRASTRUCTURE ONLY).

What runs is the UNMODIFIED reference: ``NomicBertModel`` from models/huggingface/modeling_hf_nomic_bert.py:1650 (its
pure-PyTorch tower: SDPA attention :1285-1414, gated MLP :1031-1071, post-norm blocks) and ``grad_cache_loss`` /
``clip_loss`` from loss.py:76-213, imported through ``oracle.ref_loader``.  The reference's flash-attn BiEncoder cannot be
imported in this image (``dropout_layer_norm`` / ``fused_dense_lib`` extensions are not installed), so the
three small modules around the trunk are restated here, each citing the lines it follows:

  * ``MeanPooling``        models/biencoder/modeling_biencoder.py:79-90
  * ``LogitScale``         models/biencoder/modeling_biencoder.py:30-41
  * ``BiEncoder.forward``  models/biencoder/modeling_biencoder.py:287-319 (trunk -> pool -> cast to trunk dtype -> F.normalize)

``bench.py`` uses this for ``--impl reference`` / ``cpu_baseline`` (CPU, fp32) and ``gpu_baseline`` (the same code on the
B200 under bf16 autocast: the "reference build on the same box" anchor north_star names).
"""
from __future__ import annotations

from contextlib import nullcontext

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import ref_loader


def hf_config(ref, vocab_size=30528, n_embd=768, n_head=12, n_inner=3072, n_layer=12, n_positions=512, rotary_emb_base=1000.0,
              layer_norm_epsilon=1e-12):
    """nomic-bert-base as the reference configures it for contrastive training (configs/train/mlm.yaml:33-47: SwiGLU, full
    rotary, no linear biases, post-norm; dropouts 0 as in the bench's synthetic config)."""
    return ref.hf_cfg.NomicBertConfig(
        vocab_size=vocab_size, n_embd=n_embd, n_head=n_head, n_inner=n_inner, n_layer=n_layer, n_positions=n_positions,
        activation_function="swiglu", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=layer_norm_epsilon,
        rotary_emb_fraction=1.0, rotary_emb_base=rotary_emb_base, qkv_proj_bias=False, mlp_fc1_bias=False, mlp_fc2_bias=False,
        prenorm=False, type_vocab_size=2, pad_token_id=None, rotary_scaling_factor=None)


class LogitScale(nn.Module):
    def __init__(self, logit_scale=50.0, trainable=False):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * torch.log(torch.tensor(float(logit_scale))), requires_grad=trainable)

    def forward(self, x):
        return x * self.logit_scale.exp()


class RefBiEncoder(nn.Module):
    """The reference's HF trunk + mean pooling + normalize, callable(**chunk) -> {"embedding"} with ``no_sync`` / ``training``
    as ``grad_cache_loss`` expects of a tower (loss.py:135-161)."""

    def __init__(self, ref, cfg):
        super().__init__()
        self.trunk = ref.hf.NomicBertModel(cfg, add_pooling_layer=False)

    def no_sync(self):
        return nullcontext()

    def forward(self, input_ids, attention_mask=None, normalize=True, **kw):
        hidden = self.trunk(input_ids, attention_mask=attention_mask).last_hidden_state
        if attention_mask is None:
            pooled = hidden.mean(dim=1)
        else:  # MeanPooling (:79-90): masked sum / clamp(mask sum)
            m = attention_mask.unsqueeze(-1).expand(hidden.size()).float()
            pooled = torch.sum(hidden * m, 1) / torch.clamp(m.sum(1), min=1e-9)
        pooled = pooled.to(hidden.dtype)  # :309-310
        return {"embedding": F.normalize(pooled, dim=-1) if normalize else pooled}


def build(device, **cfg_kw):
    ref = ref_loader.load()
    model = RefBiEncoder(ref, hf_config(ref, **cfg_kw)).to(device)
    return ref, model
