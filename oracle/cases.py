"""Seeded case definitions shared by the golden generator and the tests (TEST INFRASTRUCTURE ONLY).

Inputs are regenerated from numpy's frozen ``RandomState`` stream, so fixtures only hold outputs.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle.encoder import EncoderConfig

# clip_loss cases: n = local queries per rank, neg = documents per query (label stride, loss.py:115-117)
INFONCE_CASES = {
    "ws1_square": dict(ws=1, n=8, neg=1, d=32, scale=50.0, seed=11),
    "ws1_hardneg3": dict(ws=1, n=6, neg=3, d=16, scale=20.0, seed=12),
    "ws1_bidir": dict(ws=1, n=8, neg=1, d=32, scale=30.0, seed=13, bidirectional=True),
    "ws1_bidir_bad": dict(ws=1, n=4, neg=2, d=8, scale=30.0, seed=14, bidirectional=True),
    "ws2_square": dict(ws=2, n=8, neg=1, d=32, scale=50.0, seed=15),
    "ws2_hardneg2": dict(ws=2, n=5, neg=2, d=24, scale=50.0, seed=16),
    "ws1_128": dict(ws=1, n=128, neg=1, d=64, scale=50.0, seed=17),
}
DUAL_CASES = {
    "ws1": dict(ws=1, n=8, neg=1, d=32, scale=1.0 / 0.07, seed=21),
    "ws2": dict(ws=2, n=6, neg=1, d=32, scale=1.0 / 0.07, seed=22),
}
MATRYOSHKA_CASES = {
    "ws1": dict(ws=1, n=8, neg=1, d=64, scale=50.0, seed=31, dims=[64, 32, 16], weights=[1.0, 1.0, 0.5]),
    "ws2": dict(ws=2, n=8, neg=1, d=64, scale=50.0, seed=32, dims=[64, 48, 16], weights=[1.0, 0.5, 0.25]),
}
ENCODER_CASES = {
    "tiny": dict(vocab=256, n_embd=128, n_head=2, n_inner=256, n_layer=2, seq=48, batch=3, wseed=5, seed=41,
                 rope_base=1000.0, lens=[48, 17, 33]),
    "tiny3": dict(vocab=512, n_embd=192, n_head=3, n_inner=512, n_layer=3, seq=130, batch=2, wseed=6, seed=42,
                  rope_base=10000.0, lens=[130, 77]),
}
# BASELINE shape: nomic-bert-base, inputs as the reference's own parity test draws them (tests/test_flash_bert.py:52-57:
# batch 4, seqlen 512, ragged lengths in [256, 512], seed 0)
ENCODER_BASE_CASE = dict(vocab=30528, n_embd=768, n_head=12, n_inner=3072, n_layer=12, seq=512, batch=4, wseed=0, seed=0,
                         rope_base=1000.0)
VIT_CASES = {
    "tiny": dict(n_embd=128, n_head=2, n_inner=256, n_layer=2, img=64, patch=16, act="quick_gelu", batch=3, wseed=8, seed=61),
    "tiny_gelu": dict(n_embd=192, n_head=3, n_inner=384, n_layer=2, img=96, patch=32, act="gelu", batch=2, wseed=9, seed=62),
}
# real dims of BASELINE configs[2] / configs[4]: CLIP ViT-B/16 (197 tokens x 768, 12 layers) and ViT-L/14 (257 tokens x 1024, 24
# layers, 16 heads, inner 4096, patch 14 -> K = 588 on the patch GEMM); 2 images keep the fp32 CPU oracle in seconds
VIT_FULL_CASES = {
    "vit_b16": dict(n_embd=768, n_head=12, n_inner=3072, n_layer=12, img=224, patch=16, act="quick_gelu", batch=2, wseed=18, seed=71),
    "vit_l14": dict(n_embd=1024, n_head=16, n_inner=4096, n_layer=24, img=224, patch=14, act="quick_gelu", batch=2, wseed=19, seed=72),
}
GRADCACHE_CASE = dict(n=8, chunk=3, din=16, dout=32, scale=20.0, seed=51)
# unsaturated variant (loss ~ 1, not 1e-3) whose tower computes in fp32 even under autocast, so the GPU run differs from the
# reference's fp32 CPU run only by the bf16 rounding of the embeddings entering the fused loss: a tight driver check
GRADCACHE_SOFT_CASE = dict(n=8, chunk=3, din=16, dout=32, scale=4.0, seed=52, fp32=True)


def make_infonce_inputs(case):
    """Per-rank L2-normalised query shards [n,d] and document shards [n*neg,d] (float32)."""
    rs = np.random.RandomState(case["seed"])
    qs, ds = [], []
    for _ in range(case["ws"]):
        q = rs.randn(case["n"], case["d"])
        d = rs.randn(case["n"] * case["neg"], case["d"])
        # make the positives (rows r*neg of d) correlated with the queries so accuracy is non-trivial
        d[:: case["neg"]] += 0.8 * q
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        qs.append(q.astype(np.float32))
        ds.append(d.astype(np.float32))
    return qs, ds


def encoder_cfg(case) -> EncoderConfig:
    return EncoderConfig(vocab_size=case["vocab"], n_embd=case["n_embd"], n_head=case["n_head"], n_inner=case["n_inner"],
                         n_layer=case["n_layer"], rotary_emb_base=case["rope_base"])


def make_encoder_inputs(case):
    """ids [B,S] int64, ragged right-padded mask [B,S] int64, upstream grad for the embedding [B,d] float32."""
    rs = np.random.RandomState(case["seed"])
    ids = rs.randint(0, case["vocab"], size=(case["batch"], case["seq"])).astype(np.int64)
    mask = np.zeros((case["batch"], case["seq"]), dtype=np.int64)
    for b, L in enumerate(case["lens"]):
        mask[b, :L] = 1
    g = rs.randn(case["batch"], case["n_embd"]).astype(np.float32)
    return ids, mask, g


def make_gradcache_inputs(case, rank):
    rs = np.random.RandomState(case["seed"] + 100 * rank)
    xq = rs.randn(case["n"], case["din"]).astype(np.float32)
    xd = (xq + 0.5 * rs.randn(case["n"], case["din"])).astype(np.float32)
    return xq, xd


class TinyTower(torch.nn.Module):
    """Stand-in tower for the GradCache driver contract (loss.py:135-161): callable(**chunk) -> {"embedding"},
    exposes ``no_sync`` and ``training``."""

    def __init__(self, case):
        super().__init__()
        rs = np.random.RandomState(case["seed"] + 7)
        self.fp32 = bool(case.get("fp32", False))
        self.fc1 = torch.nn.Linear(case["din"], case["dout"])
        self.fc2 = torch.nn.Linear(case["dout"], case["dout"], bias=False)
        with torch.no_grad():
            self.fc1.weight.copy_(torch.from_numpy(0.3 * rs.randn(case["dout"], case["din"]).astype(np.float32)))
            self.fc1.bias.copy_(torch.from_numpy(0.1 * rs.randn(case["dout"]).astype(np.float32)))
            self.fc2.weight.copy_(torch.from_numpy(0.3 * rs.randn(case["dout"], case["dout"]).astype(np.float32)))

    def no_sync(self):
        from contextlib import nullcontext
        return nullcontext()

    def forward(self, input_ids):
        if self.fp32:
            with torch.autocast(device_type=input_ids.device.type, enabled=False):
                h = torch.tanh(self.fc1(input_ids.float()))
                return {"embedding": torch.nn.functional.normalize(self.fc2(h), dim=-1)}
        h = torch.tanh(self.fc1(input_ids))
        return {"embedding": torch.nn.functional.normalize(self.fc2(h), dim=-1)}


def vit_cfg(case):
    from oracle.vit import ViTConfig
    return ViTConfig(n_embd=case["n_embd"], n_head=case["n_head"], n_inner=case["n_inner"], n_layer=case["n_layer"],
                     img_size=case["img"], patch_size=case["patch"], activation_function=case["act"])


def make_vit_inputs(case):
    rs = np.random.RandomState(case["seed"])
    px = rs.randn(case["batch"], 3, case["img"], case["img"]).astype(np.float32)
    g = rs.randn(case["batch"], case["n_embd"]).astype(np.float32)
    return px, g
