"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.gen_golden   (needs /root/reference; CPU, gloo).
The vectors are small, seeded (numpy RandomState = frozen stream) and committed; tests regenerate the
*inputs* from the same seeds and compare outputs, so nothing reads /root/reference at test time.

What is executed from the reference, untouched:
  loss.py             clip_loss, cache_loss, grad_cache_loss (whole functions)
  distributed.py      gather_with_grad (torch.distributed.nn.all_gather over gloo)
  modeling_dual_encoder.py:46-65   (loss block, exec'd from the file text with stand-in towers)
  trainers/text_text.py:352-369    (Matryoshka loop, exec'd from the file text)
  modeling_biencoder.py:79-90      (MeanPooling, exec'd from the file text)
  models/huggingface/modeling_hf_nomic_bert.py  NomicBertModel (pure torch)
  modeling_biencoder.py  MultiHeadAttentionPooling, ClsSelector (imported; flash_attn_kvpacked_func given its published definition)
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import ref_loader  # noqa: E402
from oracle.cases import (INFONCE_CASES, DUAL_CASES, MATRYOSHKA_CASES, ENCODER_CASES, GRADCACHE_CASE, GRADCACHE_SOFT_CASE,  # noqa: E402
                          make_infonce_inputs, make_encoder_inputs, make_gradcache_inputs, encoder_cfg,
                          TinyTower)
from oracle.encoder import random_state_dict  # noqa: E402


class RefLogitScale(torch.nn.Module):
    """Same arithmetic as the reference LogitScale (modeling_biencoder.py:30-41), which cannot be imported
    here because its module pulls in flash-attn extensions; x * exp(p)."""

    def __init__(self, scale, trainable=True):
        super().__init__()
        self.logit_scale = torch.nn.Parameter(torch.ones([]) * np.log(scale), requires_grad=trainable)

    def forward(self, x):
        return x * self.logit_scale.exp()


def _ref_lines(relpath, lo, hi):
    with open(os.path.join(ref_loader.REF_ROOT, relpath)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def _init_pg(rank, ws, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)


def _infonce_worker(rank, ws, port, case, out_dir):
    _init_pg(rank, ws, port)
    ref = ref_loader.load()
    qs, ds = make_infonce_inputs(case)
    q = torch.tensor(qs[rank], requires_grad=True)
    d = torch.tensor(ds[rank], requires_grad=True)
    ls = RefLogitScale(case["scale"])

    class Tracker:
        def __init__(self):
            self.logged = {}

        def log(self, m, step=None):
            self.logged.update(m)

    tr = Tracker()
    res = {}
    try:
        loss = ref.loss.clip_loss(q, d, ls, gather_enabled=ws > 1, tracker=tr, dataset="x",
                                  bidirectional=case.get("bidirectional", False))
        loss.backward()
        res = dict(loss=np.float64(loss.item()), dq=q.grad.numpy(), dd=d.grad.numpy(),
                   dlogit=np.float64(ls.logit_scale.grad.item()), accuracy=np.float64(tr.logged["accuracy/accuracy_x"]))
    except ValueError as e:  # bidirectional with M != N (loss.py:119-123)
        res = dict(error=np.array(str(e)))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def _dual_worker(rank, ws, port, case, out_dir):
    _init_pg(rank, ws, port)
    ref = ref_loader.load()
    ts, vs = make_infonce_inputs(case)  # reuse generator: "queries" = text, "documents" = vision (un-normalised)
    t = torch.tensor(ts[rank] * 3.0, requires_grad=True)
    v = torch.tensor(vs[rank] * 0.5, requires_grad=True)
    ls = RefLogitScale(case["scale"])
    src = _ref_lines("models/dual_encoder/modeling_dual_encoder.py", 46, 66)
    ns = dict(text_outputs={"embedding": t}, vision_outputs={"embedding": v}, F=F, torch=torch, dist=dist,
              gather_with_grad=ref.distributed.gather_with_grad, self=types.SimpleNamespace(logit_scale=ls))
    exec(src, ns)
    loss = ns["metrics"]["loss"]
    loss.backward()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=np.float64(loss.item()), dtext=t.grad.numpy(),
             dvision=v.grad.numpy(), dlogit=np.float64(ls.logit_scale.grad.item()))
    dist.barrier()
    dist.destroy_process_group()


def _matryoshka_worker(rank, ws, port, case, out_dir):
    _init_pg(rank, ws, port)
    ref = ref_loader.load()
    qs, ds = make_infonce_inputs(case)
    q = torch.tensor(qs[rank] * 2.0, requires_grad=True)
    d = torch.tensor(ds[rank] * 0.7, requires_grad=True)
    ls = RefLogitScale(case["scale"], trainable=False)
    src = _ref_lines("trainers/text_text.py", 349, 369)
    ns = dict(query_outputs={"embedding": q}, document_outputs={"embedding": d}, F=F, torch=torch,
              gather_with_grad=ref.distributed.gather_with_grad, clip_loss=ref.loss.clip_loss, logit_scale=ls,
              matryoshka_dims=case["dims"], matroyshka_loss_weights=case["weights"], dataset_name="x",
              self=types.SimpleNamespace(tracker=None), kwargs={})
    exec(src, ns)
    loss = ns["loss"]
    loss.backward()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=np.float64(loss.item()), dq=q.grad.numpy(), dd=d.grad.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _gradcache_worker(rank, ws, port, case, out_dir):
    _init_pg(rank, ws, port)
    ref = ref_loader.load()
    torch.manual_seed(0)
    tower = TinyTower(case)
    xq, xd = make_gradcache_inputs(case, rank)
    ls = RefLogitScale(case["scale"], trainable=False)
    loss = ref.loss.grad_cache_loss(tower, {"input_ids": torch.tensor(xq)}, tower, {"input_ids": torch.tensor(xd)},
                                    case["chunk"], ls)
    g_gc = {k: p.grad.clone().numpy() for k, p in tower.named_parameters()}
    # the plain (non-GradCache) step on the same weights: SURVEY Appendix A.10 -- they must agree
    tower.zero_grad()
    q = tower(input_ids=torch.tensor(xq))["embedding"]
    d = tower(input_ids=torch.tensor(xd))["embedding"]
    loss2 = ref.loss.clip_loss(q, d, ls, gather_enabled=ws > 1)
    loss2.backward()
    res = dict(loss=np.float64(loss.item()), loss_plain=np.float64(loss2.item()))
    for k, g in g_gc.items():
        res["gc_" + k] = g
    for k, p in tower.named_parameters():
        res["plain_" + k] = p.grad.numpy()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def _run(worker, ws, case, port):
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        if ws == 1:
            worker(0, 1, port, case, tmp)
        else:
            mp.spawn(worker, args=(ws, port, case, tmp), nprocs=ws, join=True)
        out = {}
        for r in range(ws):
            with np.load(os.path.join(tmp, f"r{r}.npz")) as z:
                for k in z.files:
                    out[f"r{r}_{k}"] = z[k]
        return out


def gen_kat():
    """tests/test_loss.py:5-17 (identity scale; the shipped call omits logit_scale and is stale)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29431"
    dist.init_process_group("gloo", rank=0, world_size=1)
    ref = ref_loader.load()
    query = torch.tensor([[1, 2], [2, 3], [3, 4]], dtype=torch.float32)
    query /= torch.norm(query, dim=1, keepdim=True)
    document = torch.tensor([[1, 2], [3, 4], [2, 3]], dtype=torch.float32)
    document /= torch.norm(document, dim=1, keepdim=True)
    loss = ref.loss.clip_loss(query, document, lambda x: x)
    sim = torch.exp(query.matmul(document.T))
    softmax = sim / sim.sum(dim=1, keepdim=True)
    naive = -torch.log(softmax[torch.arange(3), torch.arange(3)]).mean()
    assert torch.allclose(loss, naive)
    dist.destroy_process_group()
    np.savez(os.path.join(GOLDEN, "kat_test_loss.npz"), loss=np.float64(loss.item()), naive=np.float64(naive.item()))
    print("kat", loss.item())


def gen_encoder():
    ref = ref_loader.load()
    for name, case in ENCODER_CASES.items():
        cfg = encoder_cfg(case)
        hf_cfg = ref.hf_cfg.NomicBertConfig(
            vocab_size=cfg.vocab_size, n_embd=cfg.n_embd, n_head=cfg.n_head, n_inner=cfg.n_inner, n_layer=cfg.n_layer,
            n_positions=case["seq"], activation_function="swiglu", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
            layer_norm_epsilon=cfg.layer_norm_epsilon, rotary_emb_fraction=1.0, rotary_emb_base=cfg.rotary_emb_base,
            qkv_proj_bias=False, mlp_fc1_bias=False, mlp_fc2_bias=False, prenorm=False, type_vocab_size=2,
            pad_token_id=None, rotary_scaling_factor=None)
        model = ref.hf.NomicBertModel(hf_cfg, add_pooling_layer=False)
        sd = random_state_dict(cfg, seed=case["wseed"])
        missing, unexpected = model.load_state_dict(sd, strict=True), None
        model.train()  # dropout p = 0
        ids, mask, gproj = make_encoder_inputs(case)
        ids_t, mask_t = torch.tensor(ids), torch.tensor(mask)
        out = model(ids_t, attention_mask=mask_t).last_hidden_state
        ns = {"nn": torch.nn, "torch": torch}
        exec(_ref_lines("models/biencoder/modeling_biencoder.py", 79, 90), ns)
        pooled = ns["MeanPooling"]()(out, ids_t, mask_t)
        emb = F.normalize(pooled, dim=-1)
        emb_h = F.normalize(F.layer_norm(pooled, (cfg.n_embd,)), dim=-1)  # hamming=True variant (:282-285,307)
        (emb * torch.tensor(gproj)).sum().backward()
        grads = {k: p.grad.numpy() for k, p in model.named_parameters()}
        keep = ["emb_ln.weight", "emb_ln.bias", "encoder.layers.0.attn.Wqkv.weight", "encoder.layers.0.norm1.weight",
                f"encoder.layers.{cfg.n_layer - 1}.mlp.fc2.weight", f"encoder.layers.{cfg.n_layer - 1}.mlp.fc11.weight",
                "encoder.layers.0.attn.out_proj.weight", "embeddings.token_type_embeddings.weight"]
        res = dict(hidden_valid=(out * mask_t.unsqueeze(-1)).detach().numpy(), pooled=pooled.detach().numpy(),
                   embedding=emb.detach().numpy(), embedding_hamming=emb_h.detach().numpy(),
                   gsum_word=grads["embeddings.word_embeddings.weight"] @ np.linspace(-1.0, 1.0, cfg.n_embd).astype(np.float32))
        for k in keep:
            res["g_" + k] = grads[k]
        res["gnorm_all"] = np.array([np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))])
        np.savez_compressed(os.path.join(GOLDEN, f"encoder_{name}.npz"), **res)
        print("encoder", name, "emb[0,:4]", emb[0, :4].tolist())


def gen_vit():
    """transformers.CLIPVisionModel (the reference's own ViT oracle, tests/test_flash_openclip.py:20-57) on weights that
    map to our reference-named state dict through the reference's remap_state_dict_hf_clip (models/vit/clip.py:56-173)."""
    import importlib
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from oracle.cases import VIT_CASES, make_vit_inputs, vit_cfg
    from oracle.vit import random_state_dict as vit_sd
    ref_loader.load()
    ref_loader._pkg("contrastors.models.vit", os.path.join(ref_loader.REF_ROOT, "models", "vit"))
    clip = importlib.import_module("contrastors.models.vit.clip")
    for name, case in VIT_CASES.items():
        cfg = vit_cfg(case)
        sd = vit_sd(cfg, seed=case["wseed"])
        hf_cfg = CLIPVisionConfig(hidden_size=cfg.n_embd, intermediate_size=cfg.n_inner, num_hidden_layers=cfg.n_layer,
                                  num_attention_heads=cfg.n_head, image_size=cfg.img_size, patch_size=cfg.patch_size,
                                  hidden_act=cfg.activation_function, layer_norm_eps=cfg.layer_norm_epsilon)
        hf = CLIPVisionModel(hf_cfg)
        d = cfg.n_embd
        m = {"vision_model.embeddings.class_embedding": sd["embeddings.cls_token"].reshape(d),
             "vision_model.embeddings.patch_embedding.weight": sd["embeddings.proj.weight"].reshape(d, 3, cfg.patch_size, cfg.patch_size),
             "vision_model.embeddings.position_embedding.weight": sd["embeddings.pos_embed"][0],
             "vision_model.pre_layrnorm.weight": sd["prepre_layernom.weight"], "vision_model.pre_layrnorm.bias": sd["prepre_layernom.bias"],
             "vision_model.post_layernorm.weight": sd["ln_f.weight"], "vision_model.post_layernorm.bias": sd["ln_f.bias"]}
        for i in range(cfg.n_layer):
            a, b = f"vision_model.encoder.layers.{i}.", f"layers.{i}."
            wq, wk, wv = sd[b + "attn.Wqkv.weight"].chunk(3, 0)
            bq, bk, bv = sd[b + "attn.Wqkv.bias"].chunk(3, 0)
            m.update({a + "self_attn.q_proj.weight": wq, a + "self_attn.k_proj.weight": wk, a + "self_attn.v_proj.weight": wv,
                      a + "self_attn.q_proj.bias": bq, a + "self_attn.k_proj.bias": bk, a + "self_attn.v_proj.bias": bv,
                      a + "self_attn.out_proj.weight": sd[b + "attn.out_proj.weight"], a + "self_attn.out_proj.bias": sd[b + "attn.out_proj.bias"],
                      a + "layer_norm1.weight": sd[b + "norm1.weight"], a + "layer_norm1.bias": sd[b + "norm1.bias"],
                      a + "layer_norm2.weight": sd[b + "norm2.weight"], a + "layer_norm2.bias": sd[b + "norm2.bias"],
                      a + "mlp.fc1.weight": sd[b + "mlp.fc1.weight"], a + "mlp.fc1.bias": sd[b + "mlp.fc1.bias"],
                      a + "mlp.fc2.weight": sd[b + "mlp.fc2.weight"], a + "mlp.fc2.bias": sd[b + "mlp.fc2.bias"]})
        missing, unexpected = hf.load_state_dict(m, strict=False)
        assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
        # the reference's own remap must take the HF weights back to our reference-named dict
        vit_config = types.SimpleNamespace(activation_function=cfg.activation_function, n_layer=cfg.n_layer)
        back = clip.remap_state_dict_hf_clip({k: v.clone() for k, v in hf.state_dict().items()}, vit_config)
        for k, v in sd.items():
            assert torch.equal(back[k].reshape(v.shape), v), k
        px, g = make_vit_inputs(case)
        hf.train()
        out = hf(pixel_values=torch.tensor(px)).pooler_output
        (out * torch.tensor(g)).sum().backward()
        grads = {k: p.grad for k, p in hf.named_parameters()}
        res = dict(cls=out.detach().numpy(),
                   g_fc2_last=grads[f"vision_model.encoder.layers.{cfg.n_layer - 1}.mlp.fc2.weight"].numpy(),
                   g_fc1_bias0=grads["vision_model.encoder.layers.0.mlp.fc1.bias"].numpy(),
                   g_out_proj0=grads["vision_model.encoder.layers.0.self_attn.out_proj.weight"].numpy(),
                   g_qbias0=grads["vision_model.encoder.layers.0.self_attn.q_proj.bias"].numpy(),
                   g_pos=grads["vision_model.embeddings.position_embedding.weight"].numpy(),
                   g_cls=grads["vision_model.embeddings.class_embedding"].numpy(),
                   g_patch_proj=(grads["vision_model.embeddings.patch_embedding.weight"].reshape(d, -1)
                                 @ torch.linspace(-1.0, 1.0, 3 * cfg.patch_size ** 2)).numpy(),
                   g_prepre_w=grads["vision_model.pre_layrnorm.weight"].numpy(), g_lnf_b=grads["vision_model.post_layernorm.bias"].numpy())
        np.savez_compressed(os.path.join(GOLDEN, f"vit_{name}.npz"), **res)
        print("vit", name, out[0, :4].tolist())


def main():
    if not ref_loader.available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(GOLDEN, exist_ok=True)
    gen_kat()
    port = 29440
    for name, case in INFONCE_CASES.items():
        out = _run(_infonce_worker, case["ws"], case, port)
        port += 1
        np.savez_compressed(os.path.join(GOLDEN, f"infonce_{name}.npz"), **out)
        print("infonce", name, {k: (v.item() if v.ndim == 0 else v.shape) for k, v in out.items() if "loss" in k or "error" in k})
    for name, case in DUAL_CASES.items():
        out = _run(_dual_worker, case["ws"], case, port)
        port += 1
        np.savez_compressed(os.path.join(GOLDEN, f"dual_{name}.npz"), **out)
        print("dual", name, {k: v.item() for k, v in out.items() if "loss" in k})
    for name, case in MATRYOSHKA_CASES.items():
        out = _run(_matryoshka_worker, case["ws"], case, port)
        port += 1
        np.savez_compressed(os.path.join(GOLDEN, f"matryoshka_{name}.npz"), **out)
        print("matryoshka", name, {k: v.item() for k, v in out.items() if "loss" in k})
    for ws in (1, 2):
        case = dict(GRADCACHE_CASE, ws=ws)
        out = _run(_gradcache_worker, ws, case, port)
        port += 1
        np.savez_compressed(os.path.join(GOLDEN, f"gradcache_ws{ws}.npz"), **out)
        print("gradcache", ws, {k: v.item() for k, v in out.items() if "loss" in k})
    gen_encoder()
    gen_vit()


def gen_gradcache_soft():
    """round 2: the unsaturated fp32-tower GradCache fixture (does not touch the other fixtures)."""
    port = 29490
    for ws in (1, 2):
        out = _run(_gradcache_worker, ws, dict(GRADCACHE_SOFT_CASE, ws=ws), port)
        port += 1
        np.savez_compressed(os.path.join(GOLDEN, f"gradcache_soft_ws{ws}.npz"), **out)
        print("gradcache_soft", ws, {k: v.item() for k, v in out.items() if "loss" in k})


POOLER_CASES = {  # name -> (batch, tokens, width, heads, inner, activation)
    "map_gelu": (3, 10, 64, 4, 96, "gelu"),
    "map_swiglu": (2, 7, 64, 2, 128, "swiglu"),
}


def gen_poolers():
    """round 2: the reference's OWN MultiHeadAttentionPooling / ClsSelector / projection tail on CPU (modeling_biencoder.py:44-49,
    93-152, 264-267, 300-317; layers/attention.py:312-440).  One third-party kernel is substituted by its published definition:
    flash_attn_kvpacked_func(q [B,Sq,H,D], kv [B,Sk,2,H,D], softmax_scale) = softmax(q k^T * scale) v per head (flash-attn
    2.x README / flash_attn_interface.py docstring) -- everything else (latent, Wq, Wkv, out_proj, norm1, MLP, residual, the
    [:, 0] selection) is the reference's code executing."""
    mb, att = ref_loader.load_biencoder_tail()

    def kvpacked(q, kv, dropout_p=0.0, softmax_scale=None, causal=False, **_):
        assert dropout_p == 0.0 and not causal
        k, v = kv[:, :, 0], kv[:, :, 1]
        scale = softmax_scale if softmax_scale is not None else q.shape[-1] ** -0.5
        s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v.float()).to(q.dtype)
    att.flash_attn_kvpacked_func = kvpacked
    import importlib
    importlib.import_module("contrastors.layers.mlp").swiglu = None  # "fused op unavailable": GatedMLP's own y * silu(gate) branch (mlp.py:80-81)
    for name, (B, S, d, H, inner, act) in POOLER_CASES.items():
        cfg = types.SimpleNamespace(n_embd=d, n_head=H, n_inner=inner, activation_function=act, use_flash_attn=True,
                                    fused_bias_fc=False, qkv_proj_bias=True, mlp_fc1_bias=True, mlp_fc2_bias=True, causal=False,
                                    attn_pdrop=0.0, use_rms_norm=False, layer_norm_epsilon=1e-5, num_heads_kv=None)
        torch.manual_seed(11)
        pool = mb.MultiHeadAttentionPooling(cfg).float().eval()
        with torch.no_grad():
            for p in pool.parameters():
                p.copy_(torch.randn_like(p) * (0.5 if p.dim() == 1 else p.shape[-1] ** -0.5))
        rs = np.random.RandomState(5)
        hidden = torch.tensor(rs.randn(B, S, d).astype(np.float32), requires_grad=True)
        cot = torch.tensor(rs.randn(B, d).astype(np.float32))
        out = pool(hidden, None, None)
        out.backward(cot)
        res = {"hidden": hidden.detach().numpy(), "cot": cot.numpy(), "out": out.detach().numpy(), "d_hidden": hidden.grad.numpy(),
               "cls": mb.ClsSelector()(hidden.detach(), None, None).numpy(), "n_head": np.int64(H), "activation": np.array(act)}
        for k, v in pool.state_dict().items():
            res["sd." + k] = v.numpy()
        for k, p in pool.named_parameters():
            res["g." + k] = p.grad.numpy()
        # the tail of BiEncoder.forward (modeling_biencoder.py:307-317) on the pooled rows: cast to the trunk dtype, proj, normalize
        proj = torch.nn.Linear(d, 48)
        emb = out.detach()
        if emb.dtype != torch.bfloat16:
            emb = emb.to(torch.bfloat16)   # `embedding.to(trunk_output.dtype)` with a bf16 trunk
        tail = F.normalize(proj(emb.float()), dim=-1)
        res.update(proj_w=proj.weight.detach().numpy(), proj_b=proj.bias.detach().numpy(), tail=tail.detach().numpy())
        np.savez_compressed(os.path.join(GOLDEN, f"pooler_{name}.npz"), **res)
        print("pooler", name, out[0, :3].tolist())
    # the reference's own LogitScale (modeling_biencoder.py:30-41): output, parameter gradient, state-dict key, repr
    ls = mb.LogitScale(types.SimpleNamespace(logit_scale=1 / 0.07, trainable_logit_scale=True))
    x = torch.tensor(np.random.RandomState(9).randn(5, 3).astype(np.float32))
    y = ls(x)
    y.sum().backward()
    np.savez_compressed(os.path.join(GOLDEN, "logit_scale.npz"), x=x.numpy(), y=y.detach().numpy(), dp=ls.logit_scale.grad.numpy(),
                        keys=np.array(list(ls.state_dict().keys())), p=ls.logit_scale.detach().numpy(), repr=np.array(repr(ls)))
    print("logit_scale", repr(ls))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gradcache_soft":
        gen_gradcache_soft()
    elif len(sys.argv) > 1 and sys.argv[1] == "poolers":
        gen_poolers()
    else:
        main()
