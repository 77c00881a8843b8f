"""``LogitScale`` with the reference's state-dict key and arithmetic
(/root/reference/src/contrastors/models/biencoder/modeling_biencoder.py:30-41): x * exp(p), p = log(logit_scale)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LogitScale(nn.Module):
    def __init__(self, config=None, logit_scale=None, trainable_logit_scale=None):
        super().__init__()
        if config is not None:
            logit_scale = config.logit_scale if logit_scale is None else logit_scale
            trainable = getattr(config, "trainable_logit_scale", False) if trainable_logit_scale is None else trainable_logit_scale
        else:
            trainable = bool(trainable_logit_scale)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(logit_scale), requires_grad=bool(trainable))

    def forward(self, x):
        return x * self.logit_scale.exp()

    def __repr__(self):
        return f"LogitScale(logit_scale={self.logit_scale.exp().item()}, trainable={self.logit_scale.requires_grad})"
