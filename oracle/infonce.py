"""InfoNCE oracle: closed-form float64 numpy restatement of the reference loss path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line:
  * ``clip_loss``                 /root/reference/src/contrastors/loss.py:76-132
  * ``gather_with_grad``          /root/reference/src/contrastors/distributed.py:5-12
  * ``LogitScale.forward``        /root/reference/src/contrastors/models/biencoder/modeling_biencoder.py:30-41
  * ``DualEncoder.forward`` loss  /root/reference/src/contrastors/models/dual_encoder/modeling_dual_encoder.py:46-68
  * Matryoshka loop               /root/reference/src/contrastors/trainers/text_text.py:352-369

Everything is written out explicitly (logits, log-sum-exp, closed-form gradients) in
float64 so that it is an independent statement of the arithmetic, not a call into
torch autograd.  Multi-rank behaviour is modelled by passing the list of per-rank
shards: ``all_gather`` = concatenation in rank order, its backward = sum over ranks
of the slice belonging to each rank (reduce-scatter SUM).
"""
from __future__ import annotations

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (what autocast does to MMA inputs)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    u = ((u + rounding) >> 16) << 16
    out = u.astype(np.uint32).view(np.float32)
    return np.where(np.isnan(x), x, out).astype(np.float32)


def labels_for(n: int, m: int, rank: int, world_size: int) -> np.ndarray:
    """loss.py:108-117: labels = (arange(N) + rank*N) * (M // (N*ws))."""
    stride = m // (n * world_size)
    return (np.arange(n, dtype=np.int64) + rank * n) * stride


def _lse_rows(s: np.ndarray) -> np.ndarray:
    mx = s.max(axis=1, keepdims=True)
    return (mx + np.log(np.exp(s - mx).sum(axis=1, keepdims=True)))[:, 0]


def cross_entropy_rows(s: np.ndarray, labels: np.ndarray):
    """Row-wise CE pieces: returns (mean loss, lse[N], softmax[N,M])."""
    lse = _lse_rows(s)
    n = s.shape[0]
    loss = float(np.mean(lse - s[np.arange(n), labels]))
    p = np.exp(s - lse[:, None])
    return loss, lse, p


def clip_loss_fwd_bwd(query, document, scale: float, rank: int = 0, world_size: int = 1,
                      grad_out: float = 1.0, bidirectional: bool = False):
    """One rank's ``clip_loss`` on already-gathered documents (loss.py:105-132).

    query [N,D], document [M,D] (M = all gathered rows), ``scale`` = exp(p) of LogitScale.
    Returns dict(loss, lse, argmax, labels, accuracy, dq, dd, dlogit) where
      dd is the gradient w.r.t. the *gathered* document matrix (before reduce-scatter) and
      dlogit is d loss / d p  (p = log scale; LogitScale stores p, modeling_biencoder.py:33-38).
    """
    q = np.asarray(query, dtype=np.float64)
    d = np.asarray(document, dtype=np.float64)
    n, m = q.shape[0], d.shape[0]
    labels = labels_for(n, m, rank, world_size)
    s = scale * (q @ d.T)
    loss_q, lse, p = cross_entropy_rows(s, labels)
    onehot = np.zeros_like(p)
    onehot[np.arange(n), labels] = 1.0
    if bidirectional:
        # loss.py:119-123: CE(S) + CE(d q^T) with the same N labels; torch raises unless M == N.
        if m != n:
            raise ValueError(f"Expected input batch_size ({m}) to match target batch_size ({n}).")
        st = s.T
        loss_d, _, pt = cross_entropy_rows(st, labels)
        loss = loss_q + loss_d  # NOT multiplied by world size (loss.py:123)
        onehot_t = np.zeros_like(pt)
        onehot_t[np.arange(m), labels] = 1.0
        ds = grad_out * ((p - onehot) / n + ((pt - onehot_t) / m).T)
    else:
        loss = loss_q * world_size  # loss.py:125
        ds = grad_out * world_size * (p - onehot) / n
    dq = scale * (ds @ d)
    dd = scale * (ds.T @ q)
    dlogit = float(np.sum(ds * s))
    dlogit_abs = float(np.sum(np.abs(ds * s)))  # scale of the cancelling sum (tolerance reference for tests)
    argmax = s.argmax(axis=1).astype(np.int64)  # first max wins, as ATen
    return dict(loss=loss, lse=lse, argmax=argmax, labels=labels,
                accuracy=float(np.mean(argmax == labels)), dq=dq, dd=dd, dlogit=dlogit, dlogit_abs=dlogit_abs)


def clip_loss_multirank(queries, documents, scale: float, bidirectional: bool = False):
    """All ranks of ``clip_loss(..., gather_enabled=True)`` (loss.py:100-101 + distributed.py:5-12).

    queries / documents: lists (len = world size) of per-rank shards.  Returns a list of per-rank
    dicts; ``dd`` there is the gradient of the *local* document shard after the autograd
    all-gather's backward (sum over ranks of each rank's slice), given that every rank calls
    ``loss.backward()`` on its own loss (what DDP training does).
    """
    ws = len(queries)
    all_docs = np.concatenate([np.asarray(x, dtype=np.float64) for x in documents], axis=0)
    outs = [clip_loss_fwd_bwd(queries[r], all_docs, scale, r, ws, bidirectional=bidirectional) for r in range(ws)]
    dd_total = sum(o["dd"] for o in outs)
    off = 0
    for r in range(ws):
        mr = np.asarray(documents[r]).shape[0]
        outs[r]["dd_local"] = dd_total[off:off + mr]
        off += mr
    return outs


def l2_normalize(x, eps: float = 1e-12):
    """F.normalize(x, dim=-1): x / max(||x||, eps)."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.maximum(np.sqrt((x * x).sum(axis=-1, keepdims=True)), eps)
    return x / nrm


def l2_normalize_bwd(x, gy, eps: float = 1e-12):
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    nrm = np.maximum(np.sqrt((x * x).sum(axis=-1, keepdims=True)), eps)
    y = x / nrm
    return (gy - y * (gy * y).sum(axis=-1, keepdims=True)) / nrm


def matryoshka_loss_fwd_bwd(queries, all_documents, scale: float, dims, weights, rank=0, world_size=1):
    """text_text.py:352-369: sum_w w * clip_loss(normalize(q[:, :dim]), normalize(all_d[:, :dim])).

    Inputs are the *un-normalised* embeddings (model called with normalize=False, text_text.py:325).
    Returns dict(loss, per_dim=[...], dq, dd) with gradients w.r.t. the un-normalised inputs.
    """
    q = np.asarray(queries, dtype=np.float64)
    d = np.asarray(all_documents, dtype=np.float64)
    dq = np.zeros_like(q)
    dd = np.zeros_like(d)
    total = 0.0
    per_dim = []
    dlogit = 0.0
    for w, dim in zip(weights, dims):
        qn, dn = l2_normalize(q[:, :dim]), l2_normalize(d[:, :dim])
        o = clip_loss_fwd_bwd(qn, dn, scale, rank, world_size, grad_out=w)
        total += w * o["loss"]
        per_dim.append(o)
        dq[:, :dim] += l2_normalize_bwd(q[:, :dim], o["dq"])
        dd[:, :dim] += l2_normalize_bwd(d[:, :dim], o["dd"])
        dlogit += o["dlogit"]
    return dict(loss=total, per_dim=per_dim, dq=dq, dd=dd, dlogit=dlogit)


def dual_encoder_loss_fwd_bwd(text_embs, vision_embs, scale: float):
    """modeling_dual_encoder.py:46-68 for all ranks.

    text_embs / vision_embs: lists of per-rank *un-normalised* [N,D] shards.
    loss_r = (CE(scale * v_r @ all_t^T) + CE(scale * t_r @ all_v^T)) / 2 * ws, labels = arange(N) + N*r.
    Returns per-rank dicts(loss, dtext, dvision, dlogit) with grads w.r.t. the un-normalised local shards
    (through F.normalize and both autograd all-gathers).
    """
    ws = len(text_embs)
    tn = [l2_normalize(t) for t in text_embs]
    vn = [l2_normalize(v) for v in vision_embs]
    all_t = np.concatenate(tn, 0)
    all_v = np.concatenate(vn, 0)
    n = tn[0].shape[0]
    g_local_t = [np.zeros_like(t) for t in tn]
    g_local_v = [np.zeros_like(v) for v in vn]
    g_all_t = np.zeros_like(all_t)
    g_all_v = np.zeros_like(all_v)
    outs = []
    for r in range(ws):
        # each directional CE is clip_loss with stride 1, grad_out 1/2 (the "/2*ws" factor)
        oi = clip_loss_fwd_bwd(vn[r], all_t, scale, r, ws, grad_out=0.5)
        ot = clip_loss_fwd_bwd(tn[r], all_v, scale, r, ws, grad_out=0.5)
        g_local_v[r] += oi["dq"]
        g_all_t += oi["dd"]
        g_local_t[r] += ot["dq"]
        g_all_v += ot["dd"]
        outs.append(dict(loss=0.5 * (oi["loss"] + ot["loss"]), dlogit=oi["dlogit"] + ot["dlogit"],
                         argmax_image=oi["argmax"], argmax_text=ot["argmax"]))
    for r in range(ws):
        gt = g_local_t[r] + g_all_t[r * n:(r + 1) * n]
        gv = g_local_v[r] + g_all_v[r * n:(r + 1) * n]
        outs[r]["dtext"] = l2_normalize_bwd(text_embs[r], gt)
        outs[r]["dvision"] = l2_normalize_bwd(vision_embs[r], gv)
    return outs
