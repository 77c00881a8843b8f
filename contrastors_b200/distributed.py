"""Collective helpers with the reference's names and semantics (/root/reference/src/contrastors/distributed.py).

``gather_with_grad`` (reference :5-12) is an autograd-aware all-gather: forward = concatenation of every rank's shard
in rank order, backward = reduce-scatter(SUM) of the incoming gradient (what ``torch.distributed.nn.all_gather`` does
on NCCL).  Here it is one ``all_gather_into_tensor`` straight into the output buffer (no list of tensors + ``cat``
copy) and one ``reduce_scatter_tensor`` in backward, both on NCCL over NVLink; on a non-NCCL group (the gloo CPU
tests) the backward falls back to all-reduce + slice, which is the same sum.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


# bench.py sets this to a list to collect (kind, start_event, end_event) around every collective of the path, recorded on
# the stream the collective is launched on (``comm_ms`` of the bench line); None = no events, no overhead
COMM_EVENTS = None


def _timed(kind, fn):
    from . import distributed as _self  # the list is swapped at run time: always read the module attribute
    ev = _self.COMM_EVENTS
    if ev is None or not torch.cuda.is_available():
        return fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    ev.append((kind, a, b))
    return out


def _ws():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_gather_rows(t: torch.Tensor) -> torch.Tensor:
    """Plain (no autograd) all-gather along dim 0 into one contiguous buffer."""
    ws = _ws()
    if ws == 1:
        return t
    t = t.contiguous()
    out = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _timed("allgather", lambda: dist.all_gather_into_tensor(out, t))
    return out


def reduce_scatter_rows(full: torch.Tensor) -> torch.Tensor:
    """Sum over ranks, keep this rank's row block (the backward of all_gather_rows)."""
    ws = _ws()
    if ws == 1:
        return full
    full = full.contiguous()
    n = full.shape[0] // ws
    if dist.get_backend() == "nccl":
        out = torch.empty((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        _timed("reduce_scatter", lambda: dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM))
        return out
    full = full.clone()
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    r = _rank()
    return full[r * n:(r + 1) * n].clone()


class _GatherWithGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return all_gather_rows(t)

    @staticmethod
    def backward(ctx, g):
        return reduce_scatter_rows(g)


def gather_with_grad(t):
    """reference distributed.py:5-12: identity without a process group / at world size 1; 0-dim -> [1]."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    if t.ndim == 0:
        t = t.unsqueeze(0)
    return _GatherWithGrad.apply(t)


def gather(t):
    """reference distributed.py:15-29: no-grad all-gather whose own slot keeps the local (grad-carrying) tensor."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    if t.ndim == 0:
        t = t.unsqueeze(0)
    gathered = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, t)
    gathered[dist.get_rank()] = t
    return torch.cat(gathered, dim=0)


def gather_dict(d):
    return {k: gather(v) for k, v in d.items()}


def print_rank_zero(msg):
    if dist.is_initialized() and dist.get_rank() == 0:
        print(msg)
