"""Checkpoints in the reference's on-disk layout, so a run can move between the reference trainer and
this path and real nomic-embed weights drop in.

What the reference writes (``trainers/base.py:275-344``, ``trainers/text_text.py:247-271``):

    <dir>/model/config.json            BiEncoderConfig (models/biencoder/configuration_biencoder.py) via save_pretrained
    <dir>/model/model.safetensors      BiEncoder.state_dict(): ``trunk.embeddings...``, ``trunk.emb_ln.*``,
                                       ``trunk.encoder.layers.{i}.attn.Wqkv|out_proj``, ``...mlp.fc11|fc12|fc2``,
                                       ``...norm1|norm2`` (pytorch_model.bin when safe_serialization is off)
    <dir>/model/logit_scale.pt         LogitScale.state_dict() = {"logit_scale": log-space scalar}, only when trainable
    <dir>/optimizer.pt, scheduler.pt   torch optimizer / scheduler state dicts
    <dir>/random_states_{rank}.pt      {"torch", "numpy", "random", "cuda"} RNG states

The reference's BiEncoder fetches the trunk's architecture from the Hub through ``config.model_name``; there is no
network here, so ``config.json`` additionally carries the trunk's NomicBertConfig under ``trunk_config`` with the
reference's own field names (models/encoder/configuration_nomic_bert.py) -- ``PretrainedConfig`` keeps unknown keys, so
the reference still parses the file.  Weights load by the reference's key names with or without the ``trunk.`` prefix.
"""
from __future__ import annotations

import json
import os
import random
from typing import Optional

import numpy as np
import torch

from .logit_scale import LogitScale
from .models import BiEncoder, BiEncoderConfig, NomicBertConfig

WEIGHTS_SAFE = "model.safetensors"
WEIGHTS_BIN = "pytorch_model.bin"


# ---------------------------------------------------------------------------------------------------------- config
def _trunk_config_dict(c: NomicBertConfig) -> dict:
    """NomicBertConfig -> the reference's field names (GPT2Config + configuration_nomic_bert.py:7-57) for this architecture
    (configs/train/mlm.yaml:33-47: SwiGLU, full rotary, no QKV/MLP biases, LayerNorm, post-norm)."""
    return {
        "model_type": "nomic_bert", "vocab_size": c.vocab_size, "n_embd": c.n_embd, "n_head": c.n_head, "n_layer": c.n_layer,
        "n_inner": c.n_inner, "n_positions": c.max_position, "type_vocab_size": c.type_vocab_size,
        "layer_norm_epsilon": c.layer_norm_epsilon, "initializer_range": c.initializer_range, "pad_token_id": c.pad_token_id,
        "activation_function": "swiglu", "rotary_emb_fraction": 1.0, "rotary_emb_base": c.rotary_emb_base,
        "rotary_emb_interleaved": False, "rotary_emb_scale_base": None, "qkv_proj_bias": False, "mlp_fc1_bias": False,
        "mlp_fc2_bias": False, "prenorm": False, "use_rms_norm": False, "causal": False, "parallel_block": False,
        "resid_pdrop": c.resid_pdrop, "embd_pdrop": 0.0, "attn_pdrop": 0.0,
    }


_UNSUPPORTED = {"prenorm": False, "use_rms_norm": False, "causal": False, "parallel_block": False, "qkv_proj_bias": False,
                "mlp_fc1_bias": False, "mlp_fc2_bias": False, "rotary_emb_interleaved": False}


def _trunk_config_from_dict(d: dict) -> NomicBertConfig:
    for k, want in _UNSUPPORTED.items():
        if k in d and bool(d[k]) != want:
            raise NotImplementedError(f"trunk_config.{k}={d[k]!r}: this path implements the nomic-bert contrastive architecture "
                                      f"({k}={want})")
    if float(d.get("rotary_emb_fraction", 1.0)) != 1.0:
        raise NotImplementedError("partial rotary embeddings are not implemented (rotary_emb_fraction must be 1.0)")
    if d.get("activation_function", "swiglu") != "swiglu":
        raise NotImplementedError(f"activation_function={d['activation_function']!r}: the text tower is SwiGLU")
    return NomicBertConfig(
        vocab_size=int(d["vocab_size"]), n_embd=int(d["n_embd"]), n_head=int(d["n_head"]), n_inner=int(d["n_inner"]),
        n_layer=int(d["n_layer"]), type_vocab_size=int(d.get("type_vocab_size", 2)),
        rotary_emb_base=float(d.get("rotary_emb_base", 10000.0)), layer_norm_epsilon=float(d.get("layer_norm_epsilon", 1e-12)),
        initializer_range=float(d.get("initializer_range", 0.02)), pad_token_id=d.get("pad_token_id"),
        max_position=int(d.get("n_positions", 8192)), resid_pdrop=float(d.get("resid_pdrop") or 0.0))


def config_to_dict(cfg: BiEncoderConfig) -> dict:
    """The reference's BiEncoderConfig fields (+ ``trunk_config``)."""
    from .models import nomic_bert_base
    return {
        "model_type": "biencoder", "architectures": ["BiEncoder"], "model_name": cfg.model_name, "projection_dim": cfg.projection_dim,
        "logit_scale": cfg.logit_scale, "trainable_logit_scale": cfg.trainable_logit_scale, "use_fused_kernels": True,
        "pooling": cfg.pooling, "nomic_encoder": True, "freeze": cfg.freeze, "hamming": cfg.hamming, "pretrained": False,
        "gradient_checkpointing": False, "trunk_config": _trunk_config_dict(cfg.encoder or nomic_bert_base()),
    }


def config_from_dict(d: dict) -> BiEncoderConfig:
    if "trunk_config" not in d:
        raise ValueError("config.json has no 'trunk_config': the reference resolves the trunk through the Hub (model_name), which "
                         "needs network access; add the trunk's NomicBertConfig under 'trunk_config'")
    return BiEncoderConfig(model_name=d.get("model_name", "nomic-ai/nomic-bert-2048"), pooling=d.get("pooling", "mean"),
                           hamming=bool(d.get("hamming", False)), projection_dim=d.get("projection_dim"),
                           freeze=bool(d.get("freeze", False)), logit_scale=float(d.get("logit_scale", 1 / 0.07)),
                           trainable_logit_scale=bool(d.get("trainable_logit_scale", False)),
                           encoder=_trunk_config_from_dict(d["trunk_config"]))


# ---------------------------------------------------------------------------------------------------------- weights
def save_pretrained(model: BiEncoder, output_dir: str, safe_serialization: bool = True) -> None:
    """``BiEncoder.save_pretrained`` of the reference (trainers/base.py:275-285): config.json + the state dict under the
    reference's key names (fp32 master weights, contiguous CPU copies)."""
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "config.json"), "w") as f:
        json.dump(config_to_dict(model.config), f, indent=2, sort_keys=True)
    sd = {k: v.detach().to("cpu", torch.float32).contiguous().clone() for k, v in model.state_dict().items()}
    if safe_serialization:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(output_dir, WEIGHTS_SAFE), metadata={"format": "pt"})
    else:
        torch.save(sd, os.path.join(output_dir, WEIGHTS_BIN))


def read_state_dict(model_dir: str) -> dict:
    safe, binf = os.path.join(model_dir, WEIGHTS_SAFE), os.path.join(model_dir, WEIGHTS_BIN)
    if os.path.exists(safe):
        from safetensors.torch import load_file
        return load_file(safe)
    if os.path.exists(binf):
        return torch.load(binf, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"neither {WEIGHTS_SAFE} nor {WEIGHTS_BIN} in {model_dir}")


def load_weights(model: BiEncoder, sd: dict, strict: bool = True) -> None:
    """Load a reference state dict (BiEncoder keys ``trunk.*`` or bare trunk keys) into the flat master buffer."""
    trunk, head = {}, {}
    for k, v in sd.items():
        if k.startswith("trunk."):
            trunk[k[len("trunk."):]] = v
        elif k.startswith(("embeddings.", "emb_ln.", "encoder.")):
            trunk[k] = v
        else:
            head[k] = v  # proj.* / selector.* (modeling_biencoder.py:270-285)
    model.trunk.load_reference_state_dict(trunk, strict=strict)
    own = {k: p for k, p in model.named_parameters() if not k.startswith("trunk.")}
    unexpected = [k for k in head if k not in own]
    missing = [k for k in own if k not in head]
    if strict and (unexpected or missing):
        raise KeyError(f"head parameters do not match this tower: unexpected {unexpected}, missing {missing}")
    with torch.no_grad():
        for k, v in head.items():
            if k in own:
                own[k].copy_(v.to(own[k].device, own[k].dtype))


def from_pretrained(model_dir: str, device: Optional[str] = None, strict: bool = True) -> BiEncoder:
    """``BiEncoder.from_pretrained`` of the reference (trainers/text_text.py:258-260)."""
    with open(os.path.join(model_dir, "config.json")) as f:
        cfg = config_from_dict(json.load(f))
    model = BiEncoder(cfg)
    load_weights(model, read_state_dict(model_dir), strict=strict)
    return model.to(device) if device is not None else model


# ---------------------------------------------------------------------------------------------------------- logit scale
def save_logit_scale(logit_scale: LogitScale, model_dir: str) -> bool:
    """text_text.py:247-255: written only when the scale is trainable.  Returns whether a file was written."""
    if not any(p.requires_grad for p in logit_scale.parameters()):
        return False
    os.makedirs(model_dir, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in logit_scale.state_dict().items()}, os.path.join(model_dir, "logit_scale.pt"))
    return True


def load_logit_scale(logit_scale: LogitScale, model_dir: str) -> bool:
    path = os.path.join(model_dir, "logit_scale.pt")
    if not os.path.exists(path):
        return False
    logit_scale.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
    return True


# ---------------------------------------------------------------------------------------------------------- trainer state
def save_state(output_dir: str, model: BiEncoder, logit_scale: Optional[LogitScale] = None, process_index: int = 0,
               scheduler_state: Optional[dict] = None) -> None:
    """``BaseTrainer.save_state`` (trainers/base.py:316-344): model/, optimizer.pt, scheduler.pt, random_states_{rank}.pt.
    The optimizer here is the fused AdamW on the flat buffers (``FlatParamModule.fused_adamw_step``); optimizer.pt is written
    as the ``torch.optim.AdamW.state_dict()`` of the optimizer the reference's ``configure_optimizer`` builds for the same
    tower (integer ids, decay / no-decay ``param_groups``), so ``optimizer.load_state_dict`` (base.py:300-301) accepts it
    and a file written by the reference trainer loads here."""
    os.makedirs(output_dir, exist_ok=True)
    if process_index == 0:
        save_pretrained(model, os.path.join(output_dir, "model"))
        if logit_scale is not None:
            save_logit_scale(logit_scale, os.path.join(output_dir, "model"))
        torch.save(model.trunk.optimizer_state_dict(prefix="trunk."), os.path.join(output_dir, "optimizer.pt"))
        torch.save(scheduler_state or {}, os.path.join(output_dir, "scheduler.pt"))
    states = {"torch": torch.get_rng_state(), "numpy": np.random.get_state(), "random": random.getstate(),
              "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else []}
    torch.save(states, os.path.join(output_dir, f"random_states_{process_index}.pt"))


def load_state(input_dir: str, model: BiEncoder, logit_scale: Optional[LogitScale] = None, process_index: int = 0) -> dict:
    """``BaseTrainer.load_state`` (trainers/base.py:292-314) + the model weights; returns the scheduler state dict."""
    load_weights(model, read_state_dict(os.path.join(input_dir, "model")))
    if logit_scale is not None:
        load_logit_scale(logit_scale, os.path.join(input_dir, "model"))
    model.trunk.load_optimizer_state_dict(torch.load(os.path.join(input_dir, "optimizer.pt"), map_location="cpu", weights_only=True),
                                          prefix="trunk.")
    sched = torch.load(os.path.join(input_dir, "scheduler.pt"), map_location="cpu", weights_only=True)
    states = torch.load(os.path.join(input_dir, f"random_states_{process_index}.pt"), map_location="cpu", weights_only=False)
    torch.set_rng_state(states["torch"])
    np.random.set_state(states["numpy"])
    random.setstate(states["random"])
    if torch.cuda.is_available() and len(states["cuda"]):
        torch.cuda.set_rng_state_all(states["cuda"])
    return sched
