"""InfoNCE / GradCache with the reference's names and call contracts, backed by the fused sm_100a kernels.

Reference: /root/reference/src/contrastors/loss.py
  clip_loss :76-132, get_chunked_embeddings :135-146, accumulate_gradients :149-161, cache_loss :164-184,
  grad_cache_loss :187-213; Matryoshka loop trainers/text_text.py:352-369; symmetric CLIP loss
  models/dual_encoder/modeling_dual_encoder.py:46-68.

What differs from the reference is *how*, not *what*: the [N x M] logits are never materialised in HBM (tcgen05 tiles
reduced in the GEMM epilogue), the gather of document embeddings is a single bf16 all_gather_into_tensor, the backward
emits dQ / dD from the recomputed tiles, and neither the logit scale nor autograd's grad_output is read back to the
host.  Semantics kept on purpose: the ``* world_size`` factor, the label stride for in-batch
hard negatives, the ``rank * N`` label offset, per-rank accuracy, the bidirectional branch's shape error when M != N,
and the requirement of an initialised process group.
"""
from __future__ import annotations

from contextlib import nullcontext
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .distributed import _timed, all_gather_rows, gather_with_grad, reduce_scatter_rows
from .rand_state import RandContext


def _scale_tensor(logit_scale, device) -> torch.Tensor:
    """The scalar e^p a LogitScale-style callable multiplies by (modeling_biencoder.py:37-38), as a 0-dim fp32
    device tensor that stays attached to the parameter's autograd graph (works through a DDP wrapper too)."""
    with torch.autocast(device_type=device.type, enabled=False):
        s = logit_scale(torch.ones((), device=device, dtype=torch.float32))
    if not isinstance(s, torch.Tensor):
        s = torch.tensor(float(s), device=device, dtype=torch.float32)
    return s.to(torch.float32).reshape(())


@dataclass
class _NceSpec:
    label_offset: int
    label_stride: int
    mult: float                      # loss = mult * mean_i CE_i   (world size, or ws/2 for the CLIP form)
    gather: bool = False             # all-gather `document` inside (gather_with_grad semantics)
    dims: Optional[List[int]] = None  # Matryoshka prefix dims (None = full width)
    weights: Optional[List[float]] = None
    normalize: bool = False          # L2-normalise each prefix inside the kernel (rq / rd epilogue scales)
    pregathered: Optional[torch.Tensor] = None  # bf16 [ws*N, ld] documents already gathered (GradCache stream overlap)
    out: dict = field(default_factory=dict)  # per-dim stats tensors for logging (accuracy), filled by forward


def _as_bf16_rows(x: torch.Tensor):
    """bf16, row-contiguous, row stride a multiple of 8 elements (16 B, TMA requirement)."""
    if x.dtype == torch.bfloat16 and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0:
        return x
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    y, _ = ops.rows_to_bf16(x)
    return y


class _FusedInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, document, scale_t, spec: _NceSpec):
        if not (query.is_cuda and document.is_cuda):
            raise RuntimeError("contrastors_b200.clip_loss needs CUDA tensors on a B200 (there is no CPU fallback); "
                               "the CPU restatement lives in oracle/ and is test infrastructure only")
        n, width = query.shape
        if document.dim() != 2 or document.shape[1] != width:  # torch.matmul's complaint in the reference (loss.py:109)
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({n}x{width} and {document.shape[1]}x{document.shape[0]})")
        q_bf = _as_bf16_rows(query.detach())
        if spec.pregathered is not None:
            d_bf = spec.pregathered  # gathered chunk by chunk on the side stream while the encoder was still running
        else:
            d_loc = _as_bf16_rows(document.detach())
            d_bf = all_gather_rows(d_loc) if spec.gather else d_loc
        m = d_bf.shape[0]
        dims = spec.dims or [width]
        weights = spec.weights or [1.0] * len(dims)
        ws_buf = ops.infonce_workspace(n, m, width, query.device)
        scale_dev = scale_t.detach().contiguous()
        loss = torch.zeros((), device=query.device, dtype=torch.float32)
        saved_lse, saved_rq, saved_rd = [], [], []
        q32 = d32 = None
        if spec.normalize:
            # prefix norms from the same bf16-rounded rows the MMA consumes
            q32, d32 = q_bf.float(), d_bf.float()
        single_pass = (spec.normalize and len(dims) > 1 and len(dims) <= 8 and len(set(dims)) == len(dims)
                       and all(k % 64 == 0 for k in dims) and max(dims) <= width)
        if single_pass:
            # Matryoshka: ONE accumulation over K = max(dims); every prefix's statistics come from the running sum of segment
            # products held in tensor memory (cx_infonce_mat_fwd): 2 N M K FLOPs instead of 2 N M sum(dims)
            order = sorted(range(len(dims)), key=lambda i: dims[i])
            asc = [dims[i] for i in order]
            rq_all = torch.stack([ops.row_inv_norms(q32, k) for k in asc])
            rd_all = torch.stack([ops.row_inv_norms(d32, k) for k in asc])
            lse_all, arg_all, _, st_all = ops.infonce_mat_fwd(q_bf, d_bf, asc, 1.0, scale_dev, rq_all, rd_all, spec.label_offset,
                                                              spec.label_stride)
            ctx.mat = dict(asc=asc, order=order, rq=rq_all, rd=rd_all, lse=lse_all.contiguous())
            pos = {i: j for j, i in enumerate(order)}
            for i, (w, k) in enumerate(zip(weights, dims)):
                j = pos[i]
                loss = loss + st_all[j, 0] * (w * spec.mult / n)
                spec.out[k] = dict(stats=st_all[j], argmax=arg_all[j], lse=lse_all[j])
                saved_lse.append(lse_all[j])
                saved_rq.append(rq_all[j])
                saved_rd.append(rd_all[j])
        for w, k in (() if single_pass else zip(weights, dims)):
            rq = rd = None
            if spec.normalize:
                rq, rd = ops.row_inv_norms(q32, k), ops.row_inv_norms(d32, k)
            lse, argmax, label_logit, stats = ops.infonce_fwd(q_bf, d_bf, k, 1.0, scale_dev, rq, rd, spec.label_offset,
                                                              spec.label_stride, ws_buf)
            loss = loss + stats[0] * (w * spec.mult / n)
            spec.out[k] = dict(stats=stats, argmax=argmax, lse=lse)
            saved_lse.append(lse)
            saved_rq.append(rq)
            saved_rd.append(rd)
        if not single_pass:
            ctx.mat = None
        ctx.spec, ctx.dims, ctx.weights = spec, dims, weights
        ctx.n, ctx.m, ctx.width = n, m, width
        ctx.local_rows = document.shape[0]
        ctx.q_dtype, ctx.d_dtype = query.dtype, document.dtype
        ctx.ws_buf = ws_buf
        ctx.aux = (saved_lse, saved_rq, saved_rd, q32, d32)
        ctx.save_for_backward(q_bf, d_bf, scale_dev)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        q_bf, d_bf, scale_dev = ctx.saved_tensors
        spec = ctx.spec
        saved_lse, saved_rq, saved_rd, q32, d32 = ctx.aux
        n, m, width = ctx.n, ctx.m, ctx.width
        dev = q_bf.device
        coef_dev = grad_loss.detach().to(torch.float32).contiguous()
        ldw = (width + 3) // 4 * 4  # fp32 rows must be 16-byte aligned for the TMA stores
        dq = torch.empty(n, ldw, device=dev, dtype=torch.float32)[:, :width]
        dd = torch.empty(m, ldw, device=dev, dtype=torch.float32)[:, :width]
        if spec.normalize or len(ctx.dims) > 1 or ctx.dims[0] != width:
            dq.zero_()
            dd.zero_()
        dlogit = torch.zeros((), device=dev, dtype=torch.float32)
        mat = ctx.mat
        if mat is not None and 2 <= len(mat["asc"]) <= 4 and all(w_ != 0 for w_ in ctx.weights):
            # single-accumulation Matryoshka backward (cx_infonce_mat.cu): one dS-like matrix per SEGMENT of the columns, the
            # chain through F.normalize folded into one row / column scalar per prefix
            asc, order = mat["asc"], mat["order"]
            w_asc = [float(ctx.weights[i]) for i in order]
            w_max = max(abs(w_) for w_ in w_asc)
            coef = w_max * spec.mult / n
            gamma = (mat["rq"][-1].mean() * mat["rd"][-1].mean()).reshape(1)          # typical rq * rd: keeps T in fp16's range
            coef_gamma, inv_gamma = (coef_dev.reshape(1) * gamma).contiguous(), (1.0 / gamma).contiguous()
            dq_raw, dd_raw, alpha, beta = ops.infonce_mat_bwd(q_bf, d_bf, asc, [w_ / w_max for w_ in w_asc], 1.0, scale_dev, mat["rq"],
                                                              mat["rd"], spec.label_offset, spec.label_stride, mat["lse"], coef,
                                                              coef_gamma, inv_gamma, width)
            cg = coef * coef_dev
            K = asc[-1]
            seg = torch.bucketize(torch.arange(K, device=dev), torch.tensor(asc, device=dev), right=True)   # column -> segment
            A = ((mat["rq"] ** 2) * alpha).flip(0).cumsum(0).flip(0) * cg      # [P, n]: sum over the prefixes that contain the segment
            B = ((mat["rd"] ** 2) * beta).flip(0).cumsum(0).flip(0) * cg
            dq = dq_raw
            dd = dd_raw
            dq[:, :K] -= q32[:, :K] * A.t()[:, seg]
            dd[:, :K] -= d32[:, :K] * B.t()[:, seg]
            dlogit = cg * alpha.sum()
        for i, (w, k) in (() if (mat is not None and 2 <= len(mat["asc"]) <= 4 and all(w_ != 0 for w_ in ctx.weights))
                          else enumerate(zip(ctx.weights, ctx.dims))):
            stats = torch.zeros(4, device=dev, dtype=torch.float32)
            coef = w * spec.mult / n
            if not spec.normalize and k == width and len(ctx.dims) == 1:
                ops.infonce_bwd(q_bf, d_bf, k, 1.0, scale_dev, None, None, spec.label_offset, spec.label_stride,
                                saved_lse[i], coef, coef_dev, dq, dd, False, stats, ctx.ws_buf)
            else:
                ldk = (k + 3) // 4 * 4
                gq = torch.empty(n, ldk, device=dev, dtype=torch.float32)[:, :k]
                gd = torch.empty(m, ldk, device=dev, dtype=torch.float32)[:, :k]
                ops.infonce_bwd(q_bf, d_bf, k, 1.0, scale_dev, saved_rq[i], saved_rd[i], spec.label_offset,
                                spec.label_stride, saved_lse[i], coef, coef_dev, gq, gd, False, stats, ctx.ws_buf)
                if spec.normalize:  # gq / gd are gradients w.r.t. the normalised prefixes: chain through F.normalize
                    ops.l2norm_bwd(q32, gq, saved_rq[i], k, out=dq, g_prescaled=False, accumulate=True)
                    ops.l2norm_bwd(d32, gd, saved_rd[i], k, out=dd, g_prescaled=False, accumulate=True)
                else:
                    dq[:, :k] += gq
                    dd[:, :k] += gd
            dlogit = dlogit + stats[2]
        if spec.gather:
            dd = reduce_scatter_rows(dd)
        gscale = None
        if ctx.needs_input_grad[2]:
            gscale = dlogit / scale_dev  # d loss / d (e^p); autograd chains exp'() = e^p back to p
        return dq.to(ctx.q_dtype), dd.to(ctx.d_dtype), gscale, None


def _fused_infonce(query, document, logit_scale, spec: _NceSpec):
    scale_t = _scale_tensor(logit_scale, query.device)
    return _FusedInfoNCE.apply(query, document, scale_t, spec)


def clip_loss(query, document, logit_scale, step=None, gather_enabled=False, tracker=None, dataset="",
              bidirectional=False, _pregathered=None):
    """InfoNCE for N queries against M >= N documents (reference loss.py:76-132, same signature).

    ``logit_scale`` is a callable x -> x * e^p (``LogitScale`` or a DDP-wrapped one).  Returns a 0-dim fp32 tensor
    attached to autograd: loss = CE(scale * q d^T, labels) * world_size with labels = (arange(N) + rank*N) * stride.
    """
    ws = dist.get_world_size()  # raises without a process group, exactly like the reference (loss.py:117)
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = query.shape[0]
    gather = bool(gather_enabled) and ws > 1
    m = document.shape[0] * (ws if gather else 1)
    stride = m // (n * ws)
    if bidirectional:
        if m != n:  # F.cross_entropy's complaint in the reference (loss.py:119-123)
            raise ValueError(f"Expected input batch_size ({m}) to match target batch_size ({n}).")
        spec_q = _NceSpec(label_offset=rank * n, label_stride=stride, mult=1.0, gather=gather)
        spec_d = _NceSpec(label_offset=rank * n, label_stride=stride, mult=1.0, gather=False)
        loss = _fused_infonce(query, document, logit_scale, spec_q)
        doc_full = gather_with_grad(document) if gather else document
        loss = loss + _fused_infonce(doc_full, query, logit_scale, spec_d)
        spec = spec_q
    else:
        spec = _NceSpec(label_offset=rank * n, label_stride=stride, mult=float(ws), gather=gather,
                        pregathered=_pregathered if gather else None)
        loss = _fused_infonce(query, document, logit_scale, spec)
    if tracker is not None:
        # per-rank top-1 accuracy, as the reference (loss.py:127-130); the hit count came out of the same kernel
        k = query.shape[1]
        accuracy = (spec.out[k]["stats"][1] / n).detach().cpu().item()
        tracker.log({f"accuracy/accuracy_{dataset}": accuracy}, step=step)
    return loss


def matryoshka_clip_loss(queries, all_documents, logit_scale, dims, weights, tracker=None, dataset="", step=None):
    """sum_w w * clip_loss(normalize(q[:, :dim]), normalize(all_d[:, :dim])) (reference text_text.py:352-369) as one
    autograd node: every prefix reuses the same bf16 operands (K = dim prefix of the row stride) and the fused kernel's
    per-row inverse norms, so no sliced/normalised copies are materialised.  ``all_documents`` is already gathered."""
    ws = dist.get_world_size()
    rank = dist.get_rank()
    n, m = queries.shape[0], all_documents.shape[0]
    spec = _NceSpec(label_offset=rank * n, label_stride=m // (n * ws), mult=float(ws), gather=False,
                    dims=list(dims), weights=[float(w) for w in weights], normalize=True)
    loss = _fused_infonce(queries, all_documents, logit_scale, spec)
    if tracker is not None:
        for dim in dims:
            acc = (spec.out[dim]["stats"][1] / n).detach().cpu().item()
            tracker.log({f"accuracy/accuracy_{dataset}_matryoshka_{dim}": acc}, step=step)
    return loss


def symmetric_clip_loss(text_emb, vision_emb, logit_scale):
    """(CE(scale v t_all^T) + CE(scale t v_all^T)) / 2 * world_size on un-normalised tower outputs
    (reference modeling_dual_encoder.py:46-65): normalisation, both gathers and both directions run fused."""
    ws = dist.get_world_size()
    rank = dist.get_rank()
    n = vision_emb.shape[0]
    gather = ws > 1
    mk = lambda: _NceSpec(label_offset=rank * n, label_stride=1, mult=ws / 2.0, gather=gather, normalize=True)
    loss_i = _fused_infonce(vision_emb, text_emb, logit_scale, mk())
    loss_t = _fused_infonce(text_emb, vision_emb, logit_scale, mk())
    return loss_i + loss_t


# ----------------------------------------------------------------------------------------------- GradCache
def _autocast_for(tensors):
    dev = "cuda"
    for t in (tensors.values() if isinstance(tensors, dict) else tensors):
        if isinstance(t, torch.Tensor):
            dev = t.device.type
            break
    return torch.autocast(dev, dtype=torch.bfloat16)


_CHUNK_STREAMS = {}


def _chunk_streams(model, chunks):
    """Two side streams for towers that declare ``chunk_streams_ok`` (ours): GradCache chunks are independent, so
    alternating them between two streams lets one chunk's HBM-bound kernels (LayerNorm, SwiGLU backward, RoPE, ...) run
    under the other chunk's tensor-core kernels.  Generic towers keep the reference's single-stream order."""
    if not getattr(model, "chunk_streams_ok", False) or len(chunks) < 2:
        return None
    t = chunks[0].get("input_ids")
    if t is None or not t.is_cuda:
        return None
    key = (t.device.type, t.device.index)
    if key not in _CHUNK_STREAMS:
        _CHUNK_STREAMS[key] = (torch.cuda.Stream(device=t.device), torch.cuda.Stream(device=t.device))
    # state shared by both streams is brought up to date on the launching stream first (bf16 weight shadow, RoPE tables)
    trunk = getattr(model, "trunk", None)
    if trunk is not None:
        if hasattr(trunk, "shadow"):
            trunk.shadow()
        if hasattr(trunk, "rope_tables") and t.dim() == 2:
            trunk.rope_tables(t.shape[1])
    return _CHUNK_STREAMS[key]


def _retain_budget(model, chunks) -> int:
    """How many of a tower's pass-1 chunks can keep their activations for pass 2 (memory laid out for 180 GB of HBM per GPU):
    a retained chunk runs its pass-1 forward in grad mode and pass 2 back-propagates through THAT graph instead of re-running the
    forward -- same weights, same dropout masks, identical gradients, one forward less.  A nomic-bert-base chunk of 64 x 512 tokens
    holds ~12 GB, so a B200 keeps ~12 of them: 0.6 % of the step at 256 chunks per tower (N = 1), 4.7 % at 32 (N = 8).
    Towers opt in through ``saved_bytes_per_token``; CX_RETAIN_ACTIVATIONS=0 turns it off."""
    import os
    per_token = getattr(model, "saved_bytes_per_token", None)
    if per_token is None or os.environ.get("CX_RETAIN_ACTIVATIONS", "1") == "0" or not chunks:
        return 0
    if not getattr(model, "training", False) or not any(p.requires_grad for p in model.parameters()):
        return 0
    t = chunks[0].get("input_ids")
    if t is None or not t.is_cuda or t.dim() != 2:
        return 0
    per_chunk = int(per_token() * t.shape[0] * t.shape[1] * 1.15)
    free, _ = torch.cuda.mem_get_info(t.device)
    free += torch.cuda.memory_reserved(t.device) - torch.cuda.memory_allocated(t.device)
    reserve = 3 * per_chunk + (16 << 30)  # the working chunk of pass 2, allocator slack, InfoNCE workspace, NCCL buffers
    return max(0, min(len(chunks), int((free - reserve) // max(per_chunk, 1))))


def get_chunked_embeddings(model, chunks, _retain: int = 0, _retained=None):
    """Pass 1 of GradCache (reference loss.py:135-146): no-grad bf16 forwards, one RNG snapshot per chunk.  The last ``_retain``
    chunks run in grad mode and leave their graph in ``_retained[i]`` (see ``_retain_budget``)."""
    embeddings, rand_states = [], []
    streams = _chunk_streams(model, chunks)
    main = torch.cuda.current_stream() if streams else None
    if streams:
        for s in streams:
            s.wait_stream(main)
    first_kept = len(chunks) - (_retain if _retained is not None else 0)
    for i, chunk in enumerate(chunks):
        keep = i >= first_kept
        with (torch.enable_grad() if keep else torch.no_grad()):
            rand_states.append(RandContext(chunk))
            with (torch.cuda.stream(streams[i & 1]) if streams else nullcontext()):
                with _autocast_for(chunk):
                    emb = model(**chunk)
                if streams:
                    emb["embedding"].record_stream(main)
            if keep and emb["embedding"].requires_grad:
                _retained[i] = emb["embedding"]
            embeddings.append(emb["embedding"].detach())
    if streams:
        for s in streams:
            main.wait_stream(s)
    return torch.concat(embeddings, dim=0), rand_states


_COMM_STREAMS = {}
_GATHER_EVERY = 8  # chunks per document all-gather on the side stream


def _comm_stream(device):
    key = (device.type, device.index)
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


def _chunked_embeddings_with_gather(model, chunks, n_local, _retain: int = 0, _retained=None):
    """Pass 1 for the document tower with the cross-rank gather overlapped (north_star: "GradCache loop driven from CUDA
    streams so encoder compute overlaps the gather"): as each chunk's embeddings land they are cast to bf16 into this
    rank's slice and all-gathered on a side stream straight into their final rows of the InfoNCE K operand, while the
    main stream is already running the next chunk's forward.  Returns (fp32 embeddings, rand states, gathered bf16)."""
    ws, rank = dist.get_world_size(), dist.get_rank()
    embeddings, rand_states = [], []
    gathered = None
    comm = None
    row = row0 = 0
    streams = _chunk_streams(model, chunks)
    main = torch.cuda.current_stream()
    if streams:
        for s in streams:
            s.wait_stream(main)
    first_kept = len(chunks) - (_retain if _retained is not None else 0)
    for i, chunk in enumerate(chunks):
        keep = i >= first_kept
        with (torch.enable_grad() if keep else torch.no_grad()):
            rand_states.append(RandContext(chunk))
            if gathered is None:  # allocated on the main stream before any side-stream use
                dev = chunk["input_ids"].device
                comm = _comm_stream(dev)
            with (torch.cuda.stream(streams[i & 1]) if streams else nullcontext()):
                with _autocast_for(chunk):
                    emb = model(**chunk)["embedding"]
                if keep and emb.requires_grad:
                    _retained[i] = emb
                emb = emb.detach()
                e32 = emb.float().contiguous()
                b, width = e32.shape
                if gathered is None:
                    with torch.cuda.stream(main):
                        ld = (width + 7) // 8 * 8
                        gathered = torch.zeros(ws * n_local, ld, device=e32.device, dtype=torch.bfloat16)
                    torch.cuda.current_stream().wait_stream(main)
                mine = gathered[rank * n_local + row: rank * n_local + row + b]
                ops.rows_to_bf16_into(e32, mine)
                ready = torch.cuda.Event()
                ready.record()
                if streams:
                    emb.record_stream(main)
            embeddings.append(emb)
            row += b
            # one collective per _GATHER_EVERY chunks (and one for the tail): every call costs a launch, ws output views and a
            # staging copy inside the process group, and the loss only needs the rows once the LAST chunk is in
            if (i + 1) % _GATHER_EVERY == 0 or i == len(chunks) - 1:
                outs = [gathered[r * n_local + row0: r * n_local + row] for r in range(ws)]
                mine_all = gathered[rank * n_local + row0: rank * n_local + row]
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    for s_ in (streams or ()):  # chunks alternate between two streams: the group spans both
                        comm.wait_stream(s_)
                    _timed("allgather_chunk", lambda: dist.all_gather(outs, mine_all))
                row0 = row
    if streams:
        for s in streams:
            main.wait_stream(s)
    main.wait_stream(comm)
    return torch.concat(embeddings, dim=0), rand_states, gathered


def accumulate_gradients(model, inputs, cache, rand_states, router_aux_coeff, _grad_reducer=None, _retained=None):
    """Pass 2 of GradCache (reference loss.py:149-161): re-forward each chunk with its RNG replayed and back-propagate
    <embedding, cached gradient>; DDP gradient sync only on the last chunk (``_grad_reducer``: our explicit equivalent, armed
    right before the last chunk's backward so the bucketed all-reduce runs under it)."""
    length = len(inputs)
    no_sync = getattr(model, "no_sync", nullcontext)
    sync_contexts = [no_sync] * (length - 1) + [nullcontext]
    streams = _chunk_streams(model, inputs)
    main = torch.cuda.current_stream() if streams else None
    if streams:
        for s in streams:
            s.wait_stream(main)
    for i, (inp, grad, state, sync_context) in enumerate(zip(inputs, cache, rand_states, sync_contexts)):
        kept = _retained.pop(i, None) if _retained else None
        if kept is not None:  # pass 1 kept this chunk's graph: back-propagate <embedding, cached gradient> through it directly
            if _grad_reducer is not None and i == length - 1:
                _grad_reducer.arm()
            torch.autograd.backward(kept, grad.to(kept.dtype))
            continue
        with (torch.cuda.stream(streams[i & 1]) if streams else nullcontext()):
            with sync_context():
                with state:
                    with _autocast_for(inp):
                        embedding = model(**inp)
                if _grad_reducer is not None and i == length - 1:
                    _grad_reducer.arm()
                emb = embedding["embedding"]
                if "router_loss" in embedding and embedding["router_loss"] is not None:
                    surrogate = torch.dot(emb.flatten().float(), grad.flatten().float())
                    (surrogate + embedding["router_loss"] * router_aux_coeff).backward()
                else:  # d<embedding, grad>/d embedding = grad: seed the backward with it instead of forming the dot product
                    torch.autograd.backward(emb, grad.to(emb.dtype))
    if streams:
        for s in streams:
            main.wait_stream(s)


def cache_loss(tower1, tower2, query_embeddings, document_embeddings, logit_scale, bidirectional=False, _pregathered=None):
    """Loss on the cached embeddings and its gradients w.r.t. them (reference loss.py:164-184).
    Returns (dQ, dD, loss.detach()).  tower1/tower2 are unused, as in the reference (Appendix A.6)."""
    query_embs = query_embeddings.detach().requires_grad_()
    document_embs = document_embeddings.detach().requires_grad_()
    loss = clip_loss(query_embs, document_embs, logit_scale, gather_enabled=True, bidirectional=bidirectional,
                     _pregathered=None if bidirectional else _pregathered)
    loss.backward()
    return query_embs.grad, document_embs.grad, loss.detach()


def grad_cache_loss(tower1, t1_inputs, tower2, t2_inputs, chunk_size, logit_scale, bidirectional=False,
                    router_aux_coeff=False, _grad_reducers=None):
    """GradCache step (reference loss.py:187-213): chunked no-grad embeddings for both towers, the loss and its
    embedding gradients, then chunked re-forward/backward per tower; leaves gradients in ``param.grad`` and returns
    the detached loss.  Tower 2 is skipped when it is not training (reference :210)."""
    total_bs = t1_inputs["input_ids"].shape[0]
    chunked_queries, chunked_documents = [], []
    for start in range(0, total_bs, chunk_size):
        chunked_queries.append({k: v[start:start + chunk_size] for k, v in t1_inputs.items()})
        chunked_documents.append({k: v[start:start + chunk_size] for k, v in t2_inputs.items()})

    query_embs, query_rand_states = get_chunked_embeddings(tower1, chunked_queries)
    # tower 2's pass 1 comes last: as many of ITS final chunks as HBM holds keep their activations for pass 2 (none when tower 2
    # will not be back-propagated)
    retained = {}
    n_retain = _retain_budget(tower2, chunked_documents) if getattr(tower2, "training", False) else 0
    pregathered = None
    if (dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
            and t2_inputs["input_ids"].is_cuda and total_bs % chunk_size == 0):
        # the document gather runs chunk by chunk on a side stream while the next chunk's encoder forward computes
        document_embs, doc_rand_states, pregathered = _chunked_embeddings_with_gather(tower2, chunked_documents, total_bs,
                                                                                      _retain=n_retain, _retained=retained)
    else:
        document_embs, doc_rand_states = get_chunked_embeddings(tower2, chunked_documents, _retain=n_retain, _retained=retained)

    query_cache, document_cache, loss = cache_loss(tower1, tower2, query_embs, document_embs, logit_scale,
                                                   bidirectional=bidirectional, _pregathered=pregathered)

    # explicit gradient reduction (the reference's DDP hooks): a tower's reducer is armed before the step's LAST backward
    # through that tower's weights -- tower 2's last chunk when both towers share the weights
    trunk1, trunk2 = getattr(tower1, "trunk", tower1), getattr(tower2, "trunk", tower2)
    red = _grad_reducers or {}

    def _trainable(tower):  # a tower without parameters() (a bare callable) is treated as trainable, as the reference does
        params = getattr(tower, "parameters", None)
        return params is None or any(p.requires_grad for p in params())

    # LiT (BASELINE configs[4]: frozen vision tower 1 + trainable text tower 2): the reference's pass 2 back-propagates through
    # tower 1 unconditionally and dies with "does not require grad" on a fully frozen tower (its own comment at loss.py:169,
    # "not sure this works for LiT"); a frozen tower has nothing to accumulate, so its pass 2 is skipped here
    first_pass = _trainable(tower1)
    second_pass = bool(tower2.training) and _trainable(tower2)
    r1 = red.get(id(trunk1)) if not (second_pass and trunk1 is trunk2) else None
    r2 = red.get(id(trunk2)) if second_pass else None
    if first_pass:
        accumulate_gradients(tower1, chunked_queries, query_cache.split(chunk_size), query_rand_states,
                             router_aux_coeff=router_aux_coeff, _grad_reducer=r1)
    if second_pass:
        accumulate_gradients(tower2, chunked_documents, document_cache.split(chunk_size), doc_rand_states,
                             router_aux_coeff=router_aux_coeff, _grad_reducer=r2, _retained=retained)
    retained.clear()
    return loss
