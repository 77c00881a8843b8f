"""Data-parallel plumbing: one process per GPU, NCCL over NVLink (reference: DDP wrap, trainers/text_text.py:163-170).

The tower keeps all gradients in one flat fp32 buffer, so the DDP bucket machinery collapses to a single all-reduce
(AVG) launched after the last GradCache chunk of each tower.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_gradients(*towers, average=True):
    """All-reduce every distinct tower's flat gradient buffer (AVG like DDP)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    seen = set()
    for t in towers:
        trunk = getattr(t, "trunk", t)
        g = trunk.flat_grad()
        if g.data_ptr() in seen:
            continue
        seen.add(g.data_ptr())
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if average:
            g.mul_(1.0 / dist.get_world_size())


def broadcast_parameters(*towers, src=0):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in towers:
        trunk = getattr(t, "trunk", t)
        dist.broadcast(trunk._flat, src=src)
        trunk.mark_weights_updated()
