"""ViT tower (CLIP style) and DualEncoder against the ViT oracle / HF-CLIP goldens and the dual-encoder loss oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import infonce as O
from oracle.cases import ENCODER_CASES, VIT_CASES, encoder_cfg, make_encoder_inputs, make_vit_inputs, vit_cfg
from oracle.vit import random_state_dict as vit_sd, vit_forward

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def pg():
    import os
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29537")
    created = False
    if not dist.is_initialized():  # the DualEncoder loss needs a process group, exactly like the reference
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    yield
    if created:
        dist.destroy_process_group()


def _build_vit(case):
    import contrastors_b200 as cb
    ocfg = vit_cfg(case)
    cfg = cb.ViTConfig(n_embd=ocfg.n_embd, n_head=ocfg.n_head, n_inner=ocfg.n_inner, n_layer=ocfg.n_layer, img_size=ocfg.img_size,
                       patch_size=ocfg.patch_size, activation_function=ocfg.activation_function,
                       layer_norm_epsilon=ocfg.layer_norm_epsilon)
    model = cb.VisionBiEncoder(cb.VisionBiEncoderConfig(encoder=cfg)).cuda()
    sd = vit_sd(ocfg, seed=case["wseed"])
    model.trunk.load_reference_state_dict(sd)
    return model, ocfg, sd


@pytest.mark.parametrize("name", list(VIT_CASES))
def test_vit_embedding_and_grads(name):
    case = VIT_CASES[name]
    model, ocfg, sd = _build_vit(case)
    assert {k: tuple(v.shape) for k, v in model.trunk.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    px, g = make_vit_inputs(case)
    px_t, g_t = torch.tensor(px), torch.tensor(g)
    sd32 = {k: v.clone().requires_grad_() for k, v in sd.items()}
    e32 = vit_forward(sd32, ocfg, px_t)
    (e32 * g_t).sum().backward()
    sd16 = {k: v.clone().requires_grad_() for k, v in sd.items()}
    e16 = vit_forward(sd16, ocfg, px_t, dtype=torch.bfloat16).float()
    (e16 * g_t).sum().backward()
    out = model(px_t.cuda(), normalize=False)["embedding"]
    (out * g_t.cuda()).sum().backward()

    def crit(mine, ref32, ref16, what):
        err = (mine.float().cpu() - ref32).abs().max().item()
        base = (ref16.float() - ref32).abs().max().item()
        assert err <= 3.0 * base + 1e-6 * ref32.abs().max().item(), (what, err, base)

    crit(out.detach(), e32.detach(), e16.detach(), "cls embedding")
    z = golden(f"vit_{name}.npz")  # transformers.CLIPVisionModel through the reference's remap
    assert np.abs(out.detach().cpu().numpy() - z["cls"]).max() <= 3.0 * (e16.detach() - e32.detach()).abs().max().item() + 1e-6
    trunk = model.trunk
    for k in sd:
        crit(trunk.view(trunk.flat_grad(), k), sd32[k].grad, sd16[k].grad, "grad " + k)


def test_dual_encoder_loss_and_backward():
    import contrastors_b200 as cb
    tcase, vcase = ENCODER_CASES["tiny"], VIT_CASES["tiny"]
    vision, vcfg, vsd = _build_vit(vcase)
    ocfg = encoder_cfg(tcase)
    tcfg = cb.NomicBertConfig(vocab_size=ocfg.vocab_size, n_embd=ocfg.n_embd, n_head=ocfg.n_head, n_inner=ocfg.n_inner,
                              n_layer=ocfg.n_layer, rotary_emb_base=ocfg.rotary_emb_base)
    text = cb.BiEncoder(cb.BiEncoderConfig(encoder=tcfg)).cuda()
    from oracle.encoder import biencoder_forward, random_state_dict
    tsd = random_state_dict(ocfg, seed=tcase["wseed"])
    text.trunk.load_reference_state_dict(tsd)
    model = cb.DualEncoder(text, vision, logit_scale=1 / 0.07, trainable_logit_scale=True).cuda()
    ids, mask, _ = make_encoder_inputs(tcase)
    px, _ = make_vit_inputs(vcase)
    out = model({"input_ids": torch.tensor(ids).cuda(), "attention_mask": torch.tensor(mask).cuda()},
                {"input_ids": torch.tensor(px).cuda()})
    out["loss"].backward()
    # oracle: fp32 towers -> float64 symmetric loss (modeling_dual_encoder.py:46-65)
    with torch.no_grad():
        te = biencoder_forward(tsd, ocfg, torch.tensor(ids), torch.tensor(mask), normalize=False).numpy()
        ve = vit_forward(vsd, vcfg, torch.tensor(px)).numpy()
    want = O.dual_encoder_loss_fwd_bwd([te], [ve], 1 / 0.07)[0]
    assert abs(out["loss"].item() - want["loss"]) <= 2e-2 * abs(want["loss"])
    assert abs(model.logit_scale.logit_scale.grad.item() - want["dlogit"]) <= 5e-2 * abs(want["dlogit"]) + 1e-3
    assert torch.count_nonzero(vision.trunk.flat_grad()) > 0 and torch.count_nonzero(text.trunk.flat_grad()) > 0
