"""The oracle (oracle/) against the reference's golden vectors (tests/golden, made by oracle/gen_golden.py
from the unmodified reference) and the reference's one shipped known-answer test."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import infonce as O
from oracle.cases import (DUAL_CASES, ENCODER_CASES, INFONCE_CASES, MATRYOSHKA_CASES, encoder_cfg, make_encoder_inputs,
                          make_infonce_inputs)
from oracle.encoder import biencoder_forward, random_state_dict, trunk_forward

# fp32 reference vs float64 oracle
RTOL, ATOL = 2e-4, 2e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale + atol, (np.abs(a - b).max(), scale)


def test_kat_reference_test_loss():
    # /root/reference/tests/test_loss.py:5-17 with identity scale
    z = golden("kat_test_loss.npz")
    assert abs(float(z["loss"]) - 1.0940139293670654) < 1e-7
    q = np.array([[1, 2], [2, 3], [3, 4]], dtype=np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    d = np.array([[1, 2], [3, 4], [2, 3]], dtype=np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = O.clip_loss_fwd_bwd(q, d, 1.0)
    assert abs(o["loss"] - float(z["loss"])) < 1e-6
    assert abs(o["loss"] - float(z["naive"])) < 1e-6


@pytest.mark.parametrize("name", [k for k in INFONCE_CASES if k != "ws1_bidir_bad"])
def test_infonce_cases(name):
    case = INFONCE_CASES[name]
    z = golden(f"infonce_{name}.npz")
    qs, ds = make_infonce_inputs(case)
    outs = O.clip_loss_multirank(qs, ds, case["scale"], bidirectional=case.get("bidirectional", False))
    for r in range(case["ws"]):
        o = outs[r]
        close(o["loss"], z[f"r{r}_loss"], rtol=1e-4, atol=1e-6)
        close(o["dq"], z[f"r{r}_dq"])
        close(o["dd_local"], z[f"r{r}_dd"])
        close(o["dlogit"], z[f"r{r}_dlogit"], rtol=1e-3, atol=1e-6)
        assert abs(o["accuracy"] - float(z[f"r{r}_accuracy"])) < 1e-7


def test_infonce_bidirectional_shape_error():
    case = INFONCE_CASES["ws1_bidir_bad"]
    z = golden("infonce_ws1_bidir_bad.npz")
    qs, ds = make_infonce_inputs(case)
    with pytest.raises(ValueError) as e:
        O.clip_loss_multirank(qs, ds, case["scale"], bidirectional=True)
    assert str(e.value) == str(z["r0_error"])


@pytest.mark.parametrize("name", list(DUAL_CASES))
def test_dual_encoder_loss(name):
    case = DUAL_CASES[name]
    z = golden(f"dual_{name}.npz")
    ts, vs = make_infonce_inputs(case)
    ts = [t * 3.0 for t in ts]
    vs = [v * 0.5 for v in vs]
    outs = O.dual_encoder_loss_fwd_bwd(ts, vs, case["scale"])
    for r in range(case["ws"]):
        close(outs[r]["loss"], z[f"r{r}_loss"], rtol=1e-4)
        close(outs[r]["dtext"], z[f"r{r}_dtext"])
        close(outs[r]["dvision"], z[f"r{r}_dvision"])
        close(outs[r]["dlogit"], z[f"r{r}_dlogit"], rtol=1e-3)


@pytest.mark.parametrize("name", list(MATRYOSHKA_CASES))
def test_matryoshka(name):
    case = MATRYOSHKA_CASES[name]
    z = golden(f"matryoshka_{name}.npz")
    qs, ds = make_infonce_inputs(case)
    qs = [q * 2.0 for q in qs]
    ds = [d * 0.7 for d in ds]
    ws = case["ws"]
    all_d = np.concatenate(ds, 0)
    outs = [O.matryoshka_loss_fwd_bwd(qs[r], all_d, case["scale"], case["dims"], case["weights"], r, ws) for r in range(ws)]
    dd_total = sum(o["dd"] for o in outs)
    n = qs[0].shape[0]
    for r in range(ws):
        close(outs[r]["loss"], z[f"r{r}_loss"], rtol=1e-4)
        close(outs[r]["dq"], z[f"r{r}_dq"])
        close(dd_total[r * n:(r + 1) * n], z[f"r{r}_dd"])


@pytest.mark.parametrize("name", list(ENCODER_CASES))
def test_encoder(name):
    case = ENCODER_CASES[name]
    z = golden(f"encoder_{name}.npz")
    cfg = encoder_cfg(case)
    sd = {k: v.clone().requires_grad_() for k, v in random_state_dict(cfg, seed=case["wseed"]).items()}
    ids, mask, g = make_encoder_inputs(case)
    ids_t, mask_t = torch.tensor(ids), torch.tensor(mask)
    h = trunk_forward(sd, cfg, ids_t, mask_t)
    close((h * mask_t.unsqueeze(-1)).detach().numpy(), z["hidden_valid"], rtol=1e-4, atol=1e-5)
    emb = biencoder_forward(sd, cfg, ids_t, mask_t)
    close(emb.detach().numpy(), z["embedding"], rtol=1e-4, atol=1e-6)
    emb_h = biencoder_forward(sd, cfg, ids_t, mask_t, hamming=True)
    close(emb_h.detach().numpy(), z["embedding_hamming"], rtol=1e-4, atol=1e-6)
    (emb * torch.tensor(g)).sum().backward()
    for k in z.files:
        if k.startswith("g_"):
            close(sd[k[2:]].grad.numpy(), z[k], rtol=2e-3, atol=1e-7)
    close(sd["embeddings.word_embeddings.weight"].grad.numpy() @ np.linspace(-1.0, 1.0, cfg.n_embd).astype(np.float32),
          z["gsum_word"], rtol=2e-3, atol=1e-7)
    gn = np.sqrt(sum(float((v.grad.double() ** 2).sum()) for v in sd.values()))
    close(gn, float(z["gnorm_all"][0]), rtol=1e-3)


def test_gradcache_reference_equals_plain():
    # Appendix A.10: the reference ships no asserted GradCache ground truth; its own grad_cache_loss and plain
    # clip_loss().backward() agree on the same weights (pinned here from the reference run).
    for ws in (1, 2):
        z = golden(f"gradcache_ws{ws}.npz")
        for r in range(ws):
            assert abs(float(z[f"r{r}_loss"]) - float(z[f"r{r}_loss_plain"])) < 1e-6
            for k in z.files:
                if k.startswith(f"r{r}_gc_"):
                    close(z[k], z[k.replace("_gc_", "_plain_")], rtol=1e-3, atol=1e-8)


def test_bf16_round():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.14159, 1e-30, 65504.0], dtype=np.float32)
    want = torch.tensor(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(O.bf16_round(x), want)


@pytest.mark.parametrize("name", ["tiny", "tiny_gelu"])
def test_vit_oracle_vs_hf_clip_golden(name):
    """oracle/vit.py against transformers.CLIPVisionModel outputs/gradients (generated through the reference's remap)."""
    from oracle.cases import VIT_CASES, make_vit_inputs, vit_cfg
    from oracle.vit import random_state_dict as vit_sd, vit_forward
    case = VIT_CASES[name]
    cfg = vit_cfg(case)
    z = golden(f"vit_{name}.npz")
    sd = {k: v.clone().requires_grad_() for k, v in vit_sd(cfg, seed=case["wseed"]).items()}
    px, g = make_vit_inputs(case)
    out = vit_forward(sd, cfg, torch.tensor(px))
    close(out.detach().numpy(), z["cls"], rtol=1e-4, atol=1e-5)
    (out * torch.tensor(g)).sum().backward()
    L = cfg.n_layer - 1
    close(sd[f"layers.{L}.mlp.fc2.weight"].grad.numpy(), z["g_fc2_last"], rtol=2e-3, atol=1e-7)
    close(sd["layers.0.mlp.fc1.bias"].grad.numpy(), z["g_fc1_bias0"], rtol=2e-3, atol=1e-7)
    close(sd["layers.0.attn.out_proj.weight"].grad.numpy(), z["g_out_proj0"], rtol=2e-3, atol=1e-7)
    close(sd["layers.0.attn.Wqkv.bias"].grad.numpy()[:cfg.n_embd], z["g_qbias0"], rtol=2e-3, atol=1e-7)
    close(sd["embeddings.pos_embed"].grad.numpy()[0], z["g_pos"], rtol=2e-3, atol=1e-7)
    close(sd["embeddings.cls_token"].grad.numpy().reshape(-1), z["g_cls"], rtol=2e-3, atol=1e-7)
    K = 3 * cfg.patch_size ** 2
    close(sd["embeddings.proj.weight"].grad.numpy() @ np.linspace(-1.0, 1.0, K, dtype=np.float32), z["g_patch_proj"], rtol=2e-3, atol=1e-6)
    close(sd["prepre_layernom.weight"].grad.numpy(), z["g_prepre_w"], rtol=2e-3, atol=1e-7)
    close(sd["ln_f.bias"].grad.numpy(), z["g_lnf_b"], rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("name", ["map_gelu", "map_swiglu"])
def test_pooler_oracle_vs_reference_golden(name):
    """oracle/poolers.py against the reference's OWN MultiHeadAttentionPooling / ClsSelector / projection tail (modeling_biencoder.py,
    run on CPU by oracle/gen_golden.py with flash_attn_kvpacked_func given its published definition): output, input gradient and
    every parameter gradient."""
    from oracle.poolers import map_pool, project_normalize
    z = golden(f"pooler_{name}.npz")
    sd = {k[3:]: torch.tensor(z[k]).requires_grad_() for k in z.files if k.startswith("sd.")}
    hidden = torch.tensor(z["hidden"]).requires_grad_()
    out = map_pool(sd, hidden, int(z["n_head"]), activation=str(z["activation"]))
    close(out.detach().numpy(), z["out"], rtol=1e-5, atol=1e-6)
    out.backward(torch.tensor(z["cot"]))
    close(hidden.grad.numpy(), z["d_hidden"], rtol=1e-4, atol=1e-7)
    grads = [k for k in z.files if k.startswith("g.")]
    assert len(grads) == len(sd)  # every parameter of the reference module is a parameter of the restatement
    for k in grads:
        close(sd[k[2:]].grad.numpy(), z[k], rtol=1e-4, atol=1e-7)
    assert np.array_equal(hidden.detach().numpy()[:, 0], z["cls"])  # ClsSelector
    tail = project_normalize(torch.tensor(z["out"]), torch.tensor(z["proj_w"]), torch.tensor(z["proj_b"]))
    close(tail.numpy(), z["tail"], rtol=1e-5, atol=1e-6)


def test_logit_scale_vs_reference_golden():
    """contrastors_b200.LogitScale against the reference's own class (golden from oracle/gen_golden.py::gen_poolers): arithmetic,
    parameter gradient, state-dict key, repr."""
    import types

    import contrastors_b200 as cb
    z = golden("logit_scale.npz")
    ls = cb.LogitScale(types.SimpleNamespace(logit_scale=1 / 0.07, trainable_logit_scale=True))
    y = ls(torch.tensor(z["x"]))
    assert np.array_equal(y.detach().numpy(), z["y"])
    y.sum().backward()
    close(ls.logit_scale.grad.numpy(), z["dp"], rtol=1e-6, atol=0)
    assert list(ls.state_dict().keys()) == list(z["keys"]) and np.array_equal(ls.logit_scale.detach().numpy(), z["p"])
    assert repr(ls) == str(z["repr"])
