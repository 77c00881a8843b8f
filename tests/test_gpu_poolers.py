"""Poolers and the projection head against the fp32 oracle restatement of the reference's BiEncoder tail."""
import math

import pytest
import torch

from oracle import poolers as OP
from oracle.cases import ENCODER_CASES, encoder_cfg, make_encoder_inputs
from oracle.encoder import random_state_dict, trunk_forward

pytestmark = pytest.mark.gpu


def _close(a, b, rel, what):
    err = (a.float().cpu() - b.float().cpu()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= rel * ref + 1e-6, (what, err, ref)


@pytest.mark.parametrize("B,S,d,H,act", [(3, 197, 768, 12, "gelu"), (2, 50, 128, 2, "quick_gelu"), (2, 257, 256, 4, "swiglu")])
def test_map_pooling_forward_and_grads(B, S, d, H, act):
    """MultiHeadAttentionPooling (cx_attn_pool_* + tcgen05 linears) vs the oracle on the same bf16-rounded tokens: output and the
    gradients of the tokens and of every selector parameter (bf16 operands: 2^-6 of the largest entry)."""
    from contrastors_b200.poolers import MultiHeadAttentionPooling
    torch.manual_seed(11)
    pool = MultiHeadAttentionPooling(d, H, 2 * d, activation_function=act).cuda()
    with torch.no_grad():
        for p in pool.parameters():
            p.copy_(torch.randn_like(p) * (0.5 if p.dim() < 2 else 1.0 / math.sqrt(p.shape[-1])))
        pool.norm1.weight.add_(1.0)
    hidden = (torch.randn(B * S, d, device="cuda")).to(torch.bfloat16).requires_grad_()
    g = torch.randn(B, d, device="cuda")
    out = pool(hidden, B, S)
    (out * g).sum().backward()
    sd = {k: v.detach().float().cpu().requires_grad_() for k, v in pool.state_dict().items()}
    h32 = hidden.detach().float().cpu().view(B, S, d).requires_grad_()
    ref = OP.map_pool(sd, h32, H, act, pool.norm1.eps)
    (ref * g.cpu()).sum().backward()
    _close(out, ref, 2 ** -6, "map output")
    _close(hidden.grad.view(B, S, d), h32.grad, 2 ** -5, "d tokens")
    for k, p in pool.named_parameters():
        _close(p.grad, sd[k].grad, 2 ** -5, "grad " + k)


def test_text_tower_cls_pooling_and_projection():
    """BiEncoder(pooling='cls', projection_dim=64): ClsSelector + proj + normalize (modeling_biencoder.py:44-49,270-273,307-317)
    against the fp32 oracle, by the reference's 3x-bf16-error criterion on the embedding; the projection's own gradients must
    match the oracle's."""
    import contrastors_b200 as cb
    case = ENCODER_CASES["tiny3"]
    ocfg = encoder_cfg(case)
    cfg = cb.NomicBertConfig(vocab_size=ocfg.vocab_size, n_embd=ocfg.n_embd, n_head=ocfg.n_head, n_inner=ocfg.n_inner,
                             n_layer=ocfg.n_layer, rotary_emb_base=ocfg.rotary_emb_base, layer_norm_epsilon=ocfg.layer_norm_epsilon)
    model = cb.BiEncoder(cb.BiEncoderConfig(encoder=cfg, pooling="cls", projection_dim=64)).cuda()
    sd = random_state_dict(ocfg, seed=case["wseed"])
    model.trunk.load_reference_state_dict(sd)
    torch.manual_seed(3)
    with torch.no_grad():
        model.proj.weight.copy_(torch.randn(64, ocfg.n_embd) / math.sqrt(ocfg.n_embd))
        model.proj.bias.copy_(torch.randn(64) * 0.1)
    ids, mask, _ = make_encoder_inputs(case)
    ids_t, mask_t = torch.tensor(ids), torch.tensor(mask)
    g = torch.randn(ids.shape[0], 64)
    out = model(ids_t.cuda(), attention_mask=mask_t.cuda())["embedding"]
    (out * g.cuda()).sum().backward()
    w32, b32 = model.proj.weight.detach().cpu().clone().requires_grad_(), model.proj.bias.detach().cpu().clone().requires_grad_()

    def ref_run(dtype):
        h = trunk_forward(sd, ocfg, ids_t, mask_t, dtype=dtype)
        return OP.project_normalize(h[:, 0].float(), w32, b32)

    e32 = ref_run(torch.float32)
    (e32 * g).sum().backward()
    e16 = ref_run(torch.bfloat16)
    err = (out.detach().cpu() - e32.detach()).abs().max().item()
    base = (e16.detach() - e32.detach()).abs().max().item()
    assert err <= 3.0 * base + 1e-6, (err, base)
    _close(model.proj.weight.grad, w32.grad, 3e-2, "proj.weight grad")
    _close(model.proj.bias.grad, b32.grad, 3e-2, "proj.bias grad")
    assert torch.count_nonzero(model.trunk.flat_grad()) > 0


def test_vision_tower_map_pooling_trains_through_the_trunk():
    """VisionBiEncoder(pooling='map', projection_dim): the selector sees every token of the ViT trunk and its gradient flows back
    into the trunk's flat gradient buffer; one fused training step moves trunk, selector and projection."""
    import contrastors_b200 as cb
    from contrastors_b200.trainer import _ScalarAdamW
    cfg = cb.ViTConfig(n_embd=128, n_head=2, n_inner=256, n_layer=2, img_size=64, patch_size=16)
    model = cb.VisionBiEncoder(cb.VisionBiEncoderConfig(encoder=cfg, pooling="map", projection_dim=32)).cuda()
    px = torch.randn(3, 3, 64, 64, device="cuda")
    out = model(px)["embedding"]
    assert out.shape == (3, 32) and torch.allclose(out.norm(dim=-1), torch.ones(3, device="cuda"), atol=1e-3)
    (out * torch.randn_like(out)).sum().backward()
    assert torch.count_nonzero(model.trunk.flat_grad()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in model.named_parameters() if not n.startswith("trunk."))
    assert model.selector.attn.latent.grad.abs().sum() > 0
