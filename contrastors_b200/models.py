"""nomic-bert tower on the sm_100a kernels, with the reference's module/parameter names and forward contracts.

Reference call stack:
  BiEncoder.forward            models/biencoder/modeling_biencoder.py:287-319
  NomicBertModel.forward       models/encoder/modeling_nomic_bert.py:515-587
  NomicBertEncoder.forward     :307-395  (unpad -> 12 x Block -> pad)
  Block.forward (post-norm)    layers/block.py:389-463
  FlashAttention.forward       layers/attention.py:90-245   (Wqkv -> rotary -> varlen attention -> out_proj)
  GatedMLP.forward             layers/mlp.py:68-83          (fc2(fc11(x) * silu(fc12(x))))
State-dict keys are the reference's (tests/test_huggingface.py:30-34 pins them): ``embeddings.word_embeddings.weight``,
``emb_ln.weight``, ``encoder.layers.{i}.attn.Wqkv.weight``, ``...attn.out_proj.weight``, ``...mlp.fc11.weight``,
``...mlp.fc12.weight``, ``...mlp.fc2.weight``, ``...norm1/2.weight|bias``.

B200-first layout: every parameter is a view into ONE flat fp32 master buffer (2-D weights first = the AdamW decay
group, 1-D last), with a flat fp32 gradient buffer the weight-gradient GEMMs accumulate into directly (TMA reduce-add)
and a flat bf16 shadow the forward/backward GEMMs read.  The whole tower is one autograd node: forward and backward are
explicit kernel sequences over packed (unpadded) tokens; nothing runs through ATen math.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .ops import MAJOR_MN

from .flat_params import FlatParamModule


@dataclass
class NomicBertConfig:
    """The subset of the reference's NomicBertConfig (models/encoder/configuration_nomic_bert.py:4-56) the contrastive
    towers use: SwiGLU, full rotary (NeoX), no linear biases, LayerNorm, post-norm (configs/train/mlm.yaml:33-47)."""
    vocab_size: int = 30528
    n_embd: int = 768
    n_head: int = 12
    n_inner: int = 3072
    n_layer: int = 12
    type_vocab_size: int = 2
    rotary_emb_base: float = 1000.0
    layer_norm_epsilon: float = 1e-12
    initializer_range: float = 0.02
    pad_token_id: Optional[int] = None
    max_position: int = 8192
    resid_pdrop: float = 0.0   # dropout of the fused dropout-add-LayerNorm sites and emb_drop (0.0 in nomic-bert-2048;
                               # 0.1 when a config is derived from BERT, models/encoder/bert.py:20-21)

    @property
    def head_dim(self):
        return self.n_embd // self.n_head

    @property
    def hidden_size(self):
        return self.n_embd


def nomic_bert_base(**kw) -> NomicBertConfig:
    """nomic-bert-base = BERT-base dims (models/encoder/bert.py:11-50) with vocab padded to a multiple of 64."""
    return NomicBertConfig(**kw)


def _param_specs(cfg: NomicBertConfig):
    d, I = cfg.n_embd, cfg.n_inner
    two_d = [("embeddings.word_embeddings.weight", (cfg.vocab_size, d)),
             ("embeddings.token_type_embeddings.weight", (cfg.type_vocab_size, d))]
    one_d = [("emb_ln.weight", (d,)), ("emb_ln.bias", (d,))]
    for i in range(cfg.n_layer):
        p = f"encoder.layers.{i}."
        two_d += [(p + "attn.Wqkv.weight", (3 * d, d)), (p + "attn.out_proj.weight", (d, d)),
                  (p + "mlp.fc11.weight", (I, d)), (p + "mlp.fc12.weight", (I, d)),  # contiguous: [fc11; fc12] = W1
                  (p + "mlp.fc2.weight", (d, I))]
        one_d += [(p + "norm1.weight", (d,)), (p + "norm1.bias", (d,)), (p + "norm2.weight", (d,)), (p + "norm2.bias", (d,))]
    return two_d, one_d


class NomicBertModel(FlatParamModule):
    """Trunk: ids -> last hidden state over packed tokens.  Storage: see ``flat_params.FlatParamModule``."""

    def __init__(self, config: NomicBertConfig):
        super().__init__()
        self.config = config
        assert config.head_dim == 64, "the sm_100a attention kernel is specialised for head_dim 64"
        assert config.n_embd % 64 == 0 and config.n_inner % 64 == 0
        two_d, one_d = _param_specs(config)
        self._rope = None
        self._init_flat(two_d, one_d)
        self.reset_parameters()

    def _on_apply(self):
        self._rope = None

    def reset_parameters(self, seed: Optional[int] = None):
        """N(0, initializer_range) for Linear/Embedding, (1, 0) for LayerNorm (modeling_nomic_bert.py:284-292)."""
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            for name, (off, n, shape) in self._offsets.items():
                v = self._flat[off:off + n]
                if len(shape) == 2:
                    v.copy_((torch.randn(n, generator=g) * self.config.initializer_range).to(v.device))
                elif name.endswith(".weight"):
                    v.fill_(1.0)
                else:
                    v.zero_()
            if self.config.pad_token_id is not None:
                self.view(self._flat, "embeddings.word_embeddings.weight")[self.config.pad_token_id].zero_()
        self.mark_weights_updated()

    def layer_grad_slices(self):
        """[lo, hi) of each transformer layer's 2-D weights in the flat buffers (contiguous per layer): the buckets of
        ``parallel.GradientBucketReducer``."""
        out = []
        for i in range(self.config.n_layer):
            lo = self._offsets[f"encoder.layers.{i}.attn.Wqkv.weight"][0]
            off, n, _ = self._offsets[f"encoder.layers.{i}.mlp.fc2.weight"]
            out.append((lo, off + (n + 63) // 64 * 64))
        return out

    def rope_tables(self, seqlen):
        if self._rope is None or self._rope[0].shape[0] < seqlen or self._rope[0].device != self._flat.device:
            n = max(seqlen, 512)
            dim = self.config.head_dim
            inv_freq = 1.0 / (self.config.rotary_emb_base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
            freqs = torch.outer(torch.arange(n, dtype=torch.float32), inv_freq)  # fp32, as embedding.py / HF :1148-1183
            self._rope = (torch.cos(freqs).to(self._flat.device), torch.sin(freqs).to(self._flat.device),
                          inv_freq.to(self._flat.device).contiguous())
        return self._rope

    # ---------------------------------------------------------------- forward
    def forward(self, input_ids, attention_mask=None, position_ids=None, token_type_ids=None, seq_lens=None, **kwargs):
        """Returns the last hidden state re-padded to [B, S, d] (zeros at pad positions, modeling_nomic_bert.py:392-393)."""
        packed = _pack(input_ids, attention_mask, seq_lens)
        h = _TrunkFn.apply(self._flat, self, packed, None)
        B, S = input_ids.shape
        if packed.indices is None:
            return (h.view(B, S, -1),)
        out = torch.zeros(B * S, h.shape[1], device=h.device, dtype=h.dtype)
        out[packed.indices] = h
        return (out.view(B, S, -1),)


@dataclass
class _Packed:
    ids: torch.Tensor                 # [T] int64
    cu: torch.Tensor                  # [B+1] int32
    pos: torch.Tensor                 # [T] int32
    indices: Optional[torch.Tensor]   # [T] flat positions in [B*S] (None = dense)
    total: int
    nseq: int
    max_seqlen: int


def _pack(input_ids, attention_mask, seq_lens=None) -> _Packed:
    """unpad_input (flash-attn bert_padding; modeling_nomic_bert.py:333) without its device->host round trips when the
    caller provides ``seq_lens`` (a CPU tensor / list, as a data loader has for free) or no mask at all."""
    B, S = input_ids.shape
    dev = input_ids.device
    if attention_mask is None:
        cu = torch.arange(0, (B + 1) * S, S, device=dev, dtype=torch.int32)
        ids = input_ids.reshape(-1)
        total, indices = B * S, None
    else:
        if seq_lens is not None:
            lens_cpu = torch.as_tensor(seq_lens, dtype=torch.int64, device="cpu")
            total = int(lens_cpu.sum())
            cu = torch.zeros(B + 1, dtype=torch.int32)
            cu[1:] = lens_cpu.cumsum(0).to(torch.int32)
            cu = cu.to(dev, non_blocking=True)
        else:
            lens = attention_mask.sum(dim=1, dtype=torch.int32)
            cu = torch.zeros(B + 1, device=dev, dtype=torch.int32)
            cu[1:] = lens.cumsum(0)
            total = int(cu[-1].item())  # the one host sync of the varlen path (the reference has several)
        if total == B * S:
            ids, indices = input_ids.reshape(-1), None
        else:
            indices = torch.nonzero(attention_mask.reshape(-1), as_tuple=False).reshape(-1)
            ids = input_ids.reshape(-1)[indices]
    pos = ops.token_positions(cu, total)
    return _Packed(ids.contiguous(), cu, pos, indices, total, B, S)


class _TrunkFn(torch.autograd.Function):
    """One autograd node for embeddings + all blocks (+ optional pooled head).  ``flat`` is only a handle that makes
    autograd call backward; parameter gradients are accumulated straight into the model's flat gradient buffer."""

    @staticmethod
    def forward(ctx, flat, model: "NomicBertModel", packed: _Packed, head):
        cfg = model.config
        need_grad = ctx.needs_input_grad[0]  # grad mode was on at apply() time and the tower is trainable
        W = model.shadow()
        P = model._flat
        v = model.view
        H, Dh, d = cfg.n_head, cfg.head_dim, cfg.n_embd
        scale = 1.0 / math.sqrt(Dh)
        cos_t, sin_t, inv_freq = model.rope_tables(packed.max_seqlen)
        eps = cfg.layer_norm_epsilon
        # dropout: one 64-bit seed per tower call drawn from torch's CPU generator -- RandContext replays that generator
        # for GradCache's second pass, so the re-forward regenerates exactly the same keep masks (rand_state.py)
        pdrop = float(cfg.resid_pdrop) if model.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if pdrop > 0 else 0
        site = lambda k: (seed + k * 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
        ctx.drop = (pdrop, seed)
        h, st0 = ops.embed_layernorm_fwd(packed.ids, None, v(W, "embeddings.word_embeddings.weight"),
                                         v(W, "embeddings.token_type_embeddings.weight"), v(P, "emb_ln.weight"),
                                         v(P, "emb_ln.bias"), eps, p_drop=pdrop, seed=site(0))
        saved = []
        for i in range(cfg.n_layer):
            p = f"encoder.layers.{i}."
            # Wqkv GEMM with the rotation of the q / k heads in its epilogue (cos/sin evaluated there from pos * inv_freq):
            # no separate rotary pass over qkv
            qkv = ops.gemm_qkv_rope(h, v(W, p + "attn.Wqkv.weight"), packed.pos, inv_freq, 2 * d)
            attn, lse = ops.attn_fwd(qkv, packed.cu, packed.max_seqlen, H, Dh, scale)
            o = ops.gemm(attn, v(W, p + "attn.out_proj.weight"))
            h1, st1 = ops.add_layernorm_fwd(o, h, v(P, p + "norm1.weight"), v(P, p + "norm1.bias"), eps, p_drop=pdrop,
                                            seed=site(1 + 2 * i))
            w1 = _w1(model, W, i)
            if cfg.n_inner % 128 == 0:
                a, yg = ops.gemm_swiglu(h1, w1, keep_preact=need_grad)  # SwiGLU fused into the GEMM epilogue
            else:
                yg = ops.gemm(h1, w1)
                a = ops.swiglu_fwd(yg)
            m = ops.gemm(a, v(W, p + "mlp.fc2.weight"))
            h2, st2 = ops.add_layernorm_fwd(m, h1, v(P, p + "norm2.weight"), v(P, p + "norm2.bias"), eps, p_drop=pdrop,
                                            seed=site(2 + 2 * i))
            if need_grad:
                saved.append((h, qkv, attn, lse, o, st1, h1, yg, a, m, st2))
            h = h2
        ctx.model, ctx.packed, ctx.head = model, packed, head
        ctx.saved, ctx.st0 = saved, st0
        if head is None:
            return h
        if head.get("pooling", "mean") == "cls":  # ClsSelector (modeling_biencoder.py:44-49): the first token of every sequence
            pooled = h.index_select(0, packed.cu[:-1].long()).float()
        else:
            pooled = ops.mean_pool_fwd(h, packed.cu)
        emb, head_save = ops.embed_head_fwd(pooled, head["hamming"], head["normalize"])
        ctx.head_state = (pooled, head_save)
        return emb

    @staticmethod
    def backward(ctx, g_out):
        model, packed, head = ctx.model, ctx.packed, ctx.head
        cfg = model.config
        model._ensure_grad_views()
        reducer = getattr(model, "_bucket_reducer", None)
        W, P, G = model.shadow(), model._flat, model._flat_grad
        v = model.view
        H, Dh = cfg.n_head, cfg.head_dim
        scale = 1.0 / math.sqrt(Dh)
        cos_t, sin_t, inv_freq = model.rope_tables(packed.max_seqlen)
        if head is not None:
            pooled, head_save = ctx.head_state
            gp = ops.embed_head_bwd(pooled, g_out.contiguous().float(), head_save, head["hamming"], head["normalize"])
            if head.get("pooling", "mean") == "cls":
                g_a = torch.zeros(packed.total, gp.shape[1], device=gp.device, dtype=torch.bfloat16)
                g_a.index_copy_(0, packed.cu[:-1].long(), gp.to(torch.bfloat16))
            else:
                g_a = ops.mean_pool_bwd(gp, packed.cu, packed.total)
        else:
            g_a = g_out.contiguous().to(torch.bfloat16)
        g_b = None
        pdrop, seed = ctx.drop
        site = lambda k: (seed + k * 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
        for i in reversed(range(cfg.n_layer)):
            p = f"encoder.layers.{i}."
            h, qkv, attn, lse, o, st1, h1, yg, a, m, st2 = ctx.saved[i]
            ctx.saved[i] = None
            # dz2: gradient of the residual input h1; dm: gradient of the (dropped) MLP output m -- the same tensor at p = 0
            r = ops.add_layernorm_bwd(m, h1, g_a, g_b, v(P, p + "norm2.weight"), st2, v(G, p + "norm2.weight"),
                                      v(G, p + "norm2.bias"), p_drop=pdrop, seed=site(2 + 2 * i))
            dz2, dm = r if pdrop > 0 else (r, r)
            if cfg.n_inner % 256 == 0 and dm.shape[0] >= 256:  # SwiGLU backward fused into the fc2-dgrad epilogue
                dyg = ops.gemm_swiglu_bwd(dm, v(W, p + "mlp.fc2.weight"), yg)
            else:
                dyg = ops.swiglu_bwd(ops.gemm(dm, v(W, p + "mlp.fc2.weight"), b_major=MAJOR_MN), yg)
            ops.gemm(dm, a, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "mlp.fc2.weight"), accumulate=True)
            w1 = _w1(model, W, i)
            dh1 = ops.gemm(dyg, w1, b_major=MAJOR_MN)
            ops.gemm(dyg, h1, a_major=MAJOR_MN, b_major=MAJOR_MN, out=_w1(model, G, i), accumulate=True)
            r = ops.add_layernorm_bwd(o, h, dz2, dh1, v(P, p + "norm1.weight"), st1, v(G, p + "norm1.weight"),
                                      v(G, p + "norm1.bias"), p_drop=pdrop, seed=site(1 + 2 * i))
            dz1, do = r if pdrop > 0 else (r, r)
            dattn = ops.gemm(do, v(W, p + "attn.out_proj.weight"), b_major=MAJOR_MN)
            ops.gemm(do, attn, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "attn.out_proj.weight"), accumulate=True)
            dqkv = ops.attn_bwd(qkv, attn, dattn, lse, packed.cu, packed.max_seqlen, H, Dh, scale, packed.pos, cos_t, sin_t, inv_freq)
            dh = ops.gemm(dqkv, v(W, p + "attn.Wqkv.weight"), b_major=MAJOR_MN)
            ops.gemm(dqkv, h, a_major=MAJOR_MN, b_major=MAJOR_MN, out=v(G, p + "attn.Wqkv.weight"), accumulate=True)
            if reducer is not None:
                reducer.layer_done(i)  # this layer's weight gradients are final: reduce them under the remaining layers
            g_a, g_b = dz1, dh
        ops.embed_layernorm_bwd(packed.ids, None, v(W, "embeddings.word_embeddings.weight"),
                                v(W, "embeddings.token_type_embeddings.weight"), g_a, g_b, v(P, "emb_ln.weight"), ctx.st0,
                                v(G, "embeddings.word_embeddings.weight"), v(G, "embeddings.token_type_embeddings.weight"),
                                v(G, "emb_ln.weight"), v(G, "emb_ln.bias"),
                                padding_idx=-1 if cfg.pad_token_id is None else cfg.pad_token_id, p_drop=pdrop, seed=site(0))
        if reducer is not None:
            reducer.finish()  # embeddings + every 1-D parameter (the LayerNorm gradients are reduced by a kernel per layer)
        ctx.saved = None
        return None, None, None, None


def _w1(model, buf, i):
    """[fc11; fc12] as one [2I, d] matrix (the two are adjacent in the flat buffer)."""
    off, n, shape = model._offsets[f"encoder.layers.{i}.mlp.fc11.weight"]
    off2, n2, _ = model._offsets[f"encoder.layers.{i}.mlp.fc12.weight"]
    assert off2 == off + n
    return buf[off:off + n + n2].view(2 * shape[0], shape[1])


# ======================================================================================================== BiEncoder
@dataclass
class BiEncoderConfig:
    """Fields of the reference's BiEncoderConfig (models/biencoder/configuration_biencoder.py) this path reads."""
    model_name: str = "nomic-ai/nomic-bert-2048"
    pooling: str = "mean"
    hamming: bool = False
    projection_dim: Optional[int] = None
    freeze: bool = False
    logit_scale: float = 50.0
    trainable_logit_scale: bool = False
    encoder: Optional[NomicBertConfig] = None


class BiEncoder(nn.Module):
    """Tower wrapper with the reference's forward contract (modeling_biencoder.py:287-319): trunk -> mean pooling ->
    (hamming LN) -> cast to the trunk dtype -> F.normalize -> {"embedding", "router_logits", "router_loss",
    "tokens_per_expert"}.  Pooling and the head run inside the trunk's autograd node."""

    # GradCache can alternate chunks between two CUDA streams (all shared state is atomic); measured neutral on B200 --
    # the step is power-capped, overlap only lowers the SM clock -- so it stays off and kernel timings stay clean
    chunk_streams_ok = False

    def __init__(self, config: BiEncoderConfig):
        super().__init__()
        self.config = config
        if config.pooling not in ("mean", "cls"):
            raise NotImplementedError("text tower: pooling='mean' (the shipped contrastive configs) or 'cls'; 'last' belongs to the "
                                      "decoder towers and 'map' to the vision tower (VisionBiEncoder)")
        self.trunk = NomicBertModel(config.encoder or nomic_bert_base())
        # proj (modeling_biencoder.py:270-273,312): after pooling / hamming / the cast to the trunk dtype, before normalize
        self.proj = nn.Linear(self.trunk.config.n_embd, config.projection_dim) if config.projection_dim else None
        self.frozen_trunk = bool(config.freeze)
        if self.frozen_trunk:
            self.trunk.eval()
            for p in self.trunk.parameters():
                p.requires_grad = False

    @property
    def device(self):
        return self.trunk._flat.device

    def saved_bytes_per_token(self) -> int:
        """Activations one token leaves behind for the backward (bf16 h, qkv, attn, o, h1, [y|g], a, m per layer + fp32 statistics):
        GradCache keeps whole chunks of them across its two passes when HBM allows (``loss._retain_budget``)."""
        c = self.trunk.config
        per_layer = 2 * (c.n_embd * 8 + 3 * c.n_inner) + 4 * (4 + c.n_head)  # 768*8 + 3*3072 = 15360 bf16 values at nomic-bert-base
        return c.n_layer * per_layer + 4 * c.n_embd

    def no_sync(self):
        """DDP-style context for GradCache (loss.py:151-154): gradient reduction is explicit here
        (``contrastors_b200.parallel.allreduce_gradients``), so this is a no-op context."""
        from contextlib import nullcontext
        return nullcontext()

    def forward(self, input_ids, attention_mask=None, is_padded_inputs=True, normalize=True, binarize=False, seq_lens=None,
                **kwargs):
        packed = _pack(input_ids, attention_mask, seq_lens)
        head = dict(hamming=bool(self.config.hamming), normalize=bool(normalize) and not binarize and self.proj is None,
                    pooling=self.config.pooling)
        flat = self.trunk._flat
        if self.frozen_trunk:
            with torch.no_grad():
                emb = _TrunkFn.apply(flat, self.trunk, packed, head)
        else:
            if not flat.requires_grad:
                flat.requires_grad_(True)
            emb = _TrunkFn.apply(flat, self.trunk, packed, head)
        if self.proj is not None:
            from .poolers import linear
            emb = linear(emb.to(torch.bfloat16), self.proj).float()  # the reference casts to the trunk dtype first (:309-310)
            if normalize and not binarize:
                emb = torch.nn.functional.normalize(emb, dim=-1)
        if binarize:
            emb = (emb > 0).float()
        return {"embedding": emb, "router_logits": None, "router_loss": None, "tokens_per_expert": None}
