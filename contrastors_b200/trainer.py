"""Trainer-step glue with the reference's batch schema and call order.

Reference: ``TextTextTrainer._grad_cache_forward_step`` / ``_forward_step`` (trainers/text_text.py:304-378),
``BaseTrainer.training_step`` (trainers/base.py:366-393: forward, backward, clip_grad_norm_, optimizer.step, zero_grad) and
``ImageTextTrainer._forward_step`` (trainers/image_text.py:172-178).  The batch is the reference collate output:
``query_input_ids / query_attention_mask / document_input_ids / document_attention_mask`` (+ optional ``dataset_name`` and
CPU ``*_seq_lens``).  Everything below the Python loop is the sm_100a path: GradCache chunks, fused InfoNCE, flat-buffer
gradient all-reduce, fused clip + AdamW.
"""
from __future__ import annotations

from typing import Optional

import torch

from .loss import clip_loss, gather_with_grad, grad_cache_loss, matryoshka_clip_loss
import torch.distributed as dist

from .parallel import GradientBucketReducer, allreduce_gradients, allreduce_scalar_grads
from .poolers import head_parameters


def _split_batch(batch, device):
    batch = {k: v for k, v in batch.items() if k not in ("dataset_name",)}
    batch = {k: (v if (k.endswith("seq_lens") or not isinstance(v, torch.Tensor)) else v.to(device, non_blocking=True))
             for k, v in batch.items()}
    if any(k.startswith("negative_") for k in batch):
        raise NotImplementedError("Negative sampling not supported for text-text models")  # text_text.py:346-347
    q = {k.replace("query_", ""): v for k, v in batch.items() if k.startswith("query_")}
    d = {k.replace("document_", ""): v for k, v in batch.items() if k.startswith("document_")}
    return q, d


def grad_cache_forward_step(model, batch, logit_scale, chunk_size):
    """text_text.py:304-322: returns {"loss": detached loss}; gradients are left in the flat gradient buffer."""
    q, d = _split_batch(batch, model.device)
    return {"loss": grad_cache_loss(tower1=model, t1_inputs=q, tower2=model, t2_inputs=d, chunk_size=chunk_size,
                                    logit_scale=logit_scale)}


def forward_step(model, batch, logit_scale, matryoshka_dims=None, matryoshka_loss_weights=None, tracker=None, step=None):
    """text_text.py:324-378 (no GradCache): both tower passes, gather of the documents, (Matryoshka) InfoNCE."""
    dataset = batch.get("dataset_name", "")
    q, d = _split_batch(batch, model.device)
    normalize = matryoshka_dims is None
    queries = model(**q, normalize=normalize)["embedding"]
    documents = gather_with_grad(model(**d, normalize=normalize)["embedding"])
    if matryoshka_dims:
        loss = matryoshka_clip_loss(queries, documents, logit_scale, matryoshka_dims, matryoshka_loss_weights, tracker=tracker,
                                    dataset=dataset, step=step)
    else:
        loss = clip_loss(queries, documents, logit_scale, tracker=tracker, dataset=dataset, step=step)
    return {"loss": loss}


class BatchPrefetcher:
    """Input edge (reference: the trainer copies each batch on the compute stream right before the step,
    trainers/text_text.py:308,333-343).  Wraps an iterator of HOST batches (pinned tensors in the reference's collate schema) and
    keeps one batch in flight: batch i+1 is copied host->device on a side stream while step i computes; ``next()`` makes the
    compute stream wait for that copy only.  CPU-side entries (``*_seq_lens``, ``dataset_name``) pass through untouched."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            self._next = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) and not k.endswith("seq_lens") else v)
                          for k, v in host.items()}

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self._next
        for v in batch.values():
            if isinstance(v, torch.Tensor) and v.is_cuda:
                v.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
        self._preload()
        return batch


class _ScalarAdamW:
    """AdamW for the few parameters kept outside the towers' flat buffers: the 0-dim logit scale (no decay: the reference's
    no-decay group, optimizer.py:22-23) and pooler / projection weights (decay on >= 2-D, as configure_optimizer), as a handful of
    small device ops: no host sync, no torch optimizer object (whose global
    post-step hook would force a re-cast of the towers' bf16 shadows)."""

    def __init__(self, param):
        self.p, self.step = param, 0
        self.m, self.v = torch.zeros_like(param), torch.zeros_like(param)

    @torch.no_grad()
    def update(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        g = self.p.grad
        if g is None:
            return
        self.step += 1
        b1, b2 = betas
        if weight_decay:
            self.p.mul_(1.0 - lr * weight_decay)
        self.m.mul_(b1).add_(g, alpha=1 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (self.v / (1 - b2 ** self.step)).sqrt_().add_(eps)
        self.p.addcdiv_(self.m, denom, value=-lr / (1 - b1 ** self.step))
        g.zero_()


def _step_logit_scale(logit_scale, lr, betas, eps):
    """A trainable logit scale is a model of its own in the reference (DDP-wrapped, in the optimizer's no-decay group, not
    part of ``clip_gradients``: trainers/text_text.py:172-182, base.py:361-362): average its gradient over ranks and step it."""
    params = [p for p in getattr(logit_scale, "parameters", lambda: [])() if p.requires_grad]
    if not params:
        return
    allreduce_scalar_grads(logit_scale)
    opts = logit_scale.__dict__.setdefault("_cx_scalar_adamw", {})
    for p in params:
        opts.setdefault(id(p), _ScalarAdamW(p)).update(lr, betas, eps)


def _step_tower(model, grad_scale, lr, betas, eps, weight_decay, max_grad_norm):
    """Fused clip + AdamW on the tower's flat buffers, and the same step (same global clip coefficient) for the pooler / projection
    parameters that live outside them.  A frozen trunk (LiT) is skipped; its trainable head is still stepped."""
    head = [p for p in head_parameters(model) if p.grad is not None]
    trunk_trainable = any(p.requires_grad for p in model.trunk.parameters())
    extra = None
    if head:
        for p in head:
            if dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                p.grad.mul_(1.0 / dist.get_world_size())
        extra = torch.stack([p.grad.float().pow(2).sum() for p in head]).sum()
    coef = None
    if trunk_trainable:
        coef = model.trunk.fused_adamw_step(lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                                            grad_scale=grad_scale, extra_sq_norm=extra)
    elif head and max_grad_norm:
        coef = torch.clamp(max_grad_norm / (extra.sqrt() + 1e-6), max=1.0).reshape(1)
    if head:
        opts = model.__dict__.setdefault("_cx_head_adamw", {})
        with torch.no_grad():
            for p in head:
                if coef is not None:
                    p.grad.mul_(coef.to(p.grad.dtype))
                opts.setdefault(id(p), _ScalarAdamW(p)).update(lr, betas, eps, weight_decay if p.squeeze().ndim >= 2 else 0.0)


def training_step(model, batch, logit_scale, *, lr: float, chunk_size: Optional[int] = 64, betas=(0.9, 0.999), eps=1e-8,
                  weight_decay=0.01, max_grad_norm: Optional[float] = 1.0, matryoshka_dims=None, matryoshka_loss_weights=None,
                  overlap_grad_reduce: Optional[bool] = None):
    """base.py:366-393 for a BiEncoder tower on the fused path: forward (+ backward), gradient all-reduce across ranks
    (DDP's job in the reference; in layer-ordered buckets under the last backward when ``overlap_grad_reduce``), global-norm
    clip + AdamW + zero_grad in two launches with DDP's 1 / world_size folded into the step, and the trainable logit
    scale's own all-reduce + AdamW.  ``chunk_size=None`` selects the plain (non-GradCache) step."""
    model.train()
    if overlap_grad_reduce is None:  # default: one all-reduce after the last backward; CX_OVERLAP_GRAD_REDUCE=1 selects the buckets
        import os
        overlap_grad_reduce = os.environ.get("CX_OVERLAP_GRAD_REDUCE", "0") == "1"
    reducer = None
    if overlap_grad_reduce and dist.is_initialized() and dist.get_world_size() > 1 and hasattr(model.trunk, "layer_grad_slices"):
        reducer = model.__dict__.get("_cx_reducer")
        if reducer is None:
            reducer = model.__dict__["_cx_reducer"] = GradientBucketReducer(model.trunk)
    if chunk_size:
        q, d = _split_batch(batch, model.device)
        out = {"loss": grad_cache_loss(tower1=model, t1_inputs=q, tower2=model, t2_inputs=d, chunk_size=chunk_size,
                                       logit_scale=logit_scale, _grad_reducers={id(model.trunk): reducer} if reducer else None)}
    else:
        out = forward_step(model, batch, logit_scale, matryoshka_dims, matryoshka_loss_weights)
        if reducer is not None:
            reducer.arm()
        out["loss"].backward()
    grad_scale = reducer.wait() if reducer is not None else allreduce_gradients(model, average=False)
    _step_tower(model, grad_scale, lr, betas, eps, weight_decay, max_grad_norm)
    _step_logit_scale(logit_scale, lr, betas, eps)
    return out["loss"].detach()


def dual_training_step(vision, text, vision_inputs, text_inputs, logit_scale, *, lr: float, chunk_size: int = 64, betas=(0.9, 0.999),
                       eps=1e-8, weight_decay=0.01, max_grad_norm: Optional[float] = 1.0, bidirectional: Optional[bool] = None):
    """Image-text GradCache step (BASELINE configs[2] / [4]): ``grad_cache_loss(tower1 = vision, tower2 = text)`` as SURVEY section 8d
    sets it up (pixels travel under the key ``input_ids``, image_text_loader.py:339), each trainable tower's flat gradient
    all-reduced and stepped with its own fused clip + AdamW, the (trainable) logit scale stepped as in ``training_step``.  A frozen
    tower (LiT) is skipped by both the gradient pass and the optimizer.  ``bidirectional`` defaults to the reference's validity
    rule for ``clip_loss(bidirectional=True)``: only at world size 1 (loss.py:119-123 needs M == N)."""
    vision.train()
    text.train()
    ws = dist.get_world_size() if dist.is_initialized() else 1
    if bidirectional is None:
        bidirectional = ws == 1
    loss = grad_cache_loss(tower1=vision, t1_inputs=vision_inputs, tower2=text, t2_inputs=text_inputs, chunk_size=chunk_size,
                           logit_scale=logit_scale, bidirectional=bidirectional)
    for tower in (vision, text):
        scale = allreduce_gradients(tower, average=False)  # (skips a frozen trunk)
        _step_tower(tower, scale, lr, betas, eps, weight_decay, max_grad_norm)
    _step_logit_scale(logit_scale, lr, betas, eps)
    return loss.detach()
