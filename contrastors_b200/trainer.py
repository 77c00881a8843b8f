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
from .parallel import allreduce_gradients


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


def training_step(model, batch, logit_scale, *, lr: float, chunk_size: Optional[int] = 64, betas=(0.9, 0.999), eps=1e-8,
                  weight_decay=0.01, max_grad_norm: Optional[float] = 1.0, matryoshka_dims=None, matryoshka_loss_weights=None):
    """base.py:366-393 for a BiEncoder tower on the fused path: forward (+ backward), gradient all-reduce across ranks
    (DDP's job in the reference), global-norm clip + AdamW + zero_grad in two launches.  ``chunk_size=None`` selects the
    plain (non-GradCache) step."""
    model.train()
    if chunk_size:
        out = grad_cache_forward_step(model, batch, logit_scale, chunk_size)
    else:
        out = forward_step(model, batch, logit_scale, matryoshka_dims, matryoshka_loss_weights)
        out["loss"].backward()
    allreduce_gradients(model)
    model.trunk.fused_adamw_step(lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
    return out["loss"].detach()
