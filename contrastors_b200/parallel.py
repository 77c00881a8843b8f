"""Data-parallel plumbing: one process per GPU, NCCL over NVLink (reference: DDP wrap, trainers/text_text.py:163-180).

The tower keeps all gradients in one flat fp32 buffer, so the DDP bucket machinery collapses to collectives over slices
of that buffer.  Two forms:

* ``allreduce_gradients(*towers)``: one all-reduce per distinct tower after the last backward (SUM; the ``1 / world_size``
  of DDP's average is folded into the fused AdamW's ``grad_scale`` by ``trainer.training_step`` instead of an extra pass
  over 137 M gradients, or applied here with ``average=True`` for callers that use a torch optimizer).
* ``GradientBucketReducer``: the same reduction cut into layer-ordered buckets that are launched on a communication
  stream from inside the LAST backward of the step, as soon as a layer's weight gradients are final, so the transfer runs
  under the remaining layers' backward kernels (the reference gets this from DDP's bucket hooks, fired under
  ``accumulate_gradients``'s last chunk, loss.py:151-154).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .distributed import _timed


def _distinct_trunks(towers):
    seen, out = set(), []
    for t in towers:
        trunk = getattr(t, "trunk", t)
        g = trunk.flat_grad()
        params = getattr(trunk, "parameters", None)
        if g.data_ptr() in seen or (params is not None and not any(p.requires_grad for p in params())):
            continue  # the same weights twice (tower1 is tower2) are reduced once; frozen towers (LiT) not at all
        seen.add(g.data_ptr())
        out.append(trunk)
    return out


def allreduce_gradients(*towers, average=True) -> float:
    """All-reduce (SUM) every distinct trainable tower's flat gradient buffer.  ``average=True`` scales by 1 / world_size
    like DDP; with ``average=False`` the factor is returned for the caller to fold into the optimizer step."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1.0
    ws = dist.get_world_size()
    for trunk in _distinct_trunks(towers):
        g = trunk.flat_grad()
        _timed("grad_allreduce", lambda: dist.all_reduce(g, op=dist.ReduceOp.SUM))
        if average:
            g.mul_(1.0 / ws)
    return 1.0 if average else 1.0 / ws


def allreduce_scalar_grads(module, average=True):
    """Gradient all-reduce for a small trainable module kept outside the towers (the logit scale: the reference wraps it in
    its own DDP, trainers/text_text.py:172-180)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for p in module.parameters():
        if p.requires_grad and p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            if average:
                p.grad.mul_(1.0 / dist.get_world_size())


class GradientBucketReducer:
    """Overlapped gradient reduction for one tower.  ``arm()`` before the step's last backward; the tower's backward calls
    ``layer_done(i)`` after layer i's weight gradients are accumulated (layers finish in the order n_layer-1 .. 0) and
    ``finish()`` after the embedding gradients; each call launches the all-reduce of the now-final slice on the
    communication stream.  ``wait()`` joins the stream.  Buckets: one per transformer layer's 2-D weights (contiguous in the
    flat buffer), then the embedding tables + all 1-D parameters."""

    def __init__(self, trunk):
        self.trunk = trunk
        self.armed = False
        self.stream = None
        self._layer_slices: List[Tuple[int, int]] = trunk.layer_grad_slices()
        self._done = 0

    def arm(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        if self.stream is None and self.trunk.flat_grad().is_cuda:
            self.stream = torch.cuda.Stream(device=self.trunk.flat_grad().device)
        self.armed = True
        self._done = 0
        self.trunk._bucket_reducer = self

    def _reduce(self, lo, hi):
        g = self.trunk.flat_grad()[lo:hi]
        if self.stream is None:  # host tensors (the gloo tests): same buckets, no stream
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            _timed("grad_allreduce_bucket", lambda: dist.all_reduce(g, op=dist.ReduceOp.SUM))

    def layer_done(self, i):
        if self.armed:
            lo, hi = self._layer_slices[i]
            self._reduce(lo, hi)
            self._done += 1

    def finish(self):
        """Everything not covered by a layer bucket (embeddings before the first layer slice, 1-D parameters after the last)."""
        if not self.armed:
            return
        first = min(lo for lo, _ in self._layer_slices)
        last = max(hi for _, hi in self._layer_slices)
        if first > 0:
            self._reduce(0, first)
        if last < self.trunk._n_total:
            self._reduce(last, self.trunk._n_total)
        self.armed = False
        self.trunk._bucket_reducer = None

    def wait(self) -> float:
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        ws = dist.get_world_size() if dist.is_initialized() else 1
        return 1.0 / ws


def broadcast_parameters(*towers, src=0):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in towers:
        trunk = getattr(t, "trunk", t)
        dist.broadcast(trunk._flat, src=src)
        trunk.mark_weights_updated()
