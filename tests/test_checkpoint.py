"""Checkpoint compatibility with the reference's on-disk layout (trainers/base.py:275-344,
trainers/text_text.py:247-271).  CPU only: the modules hold their parameters on the CPU here, no kernels run."""
import json
import os

import numpy as np
import pytest
import torch

import contrastors_b200 as cb
from contrastors_b200 import checkpoint as ck
from oracle import ref_loader
from oracle.cases import ENCODER_CASES, encoder_cfg
from oracle.encoder import random_state_dict


def _tiny(hamming=False, freeze=False):
    case = ENCODER_CASES["tiny"]
    ocfg = encoder_cfg(case)
    cfg = cb.NomicBertConfig(vocab_size=ocfg.vocab_size, n_embd=ocfg.n_embd, n_head=ocfg.n_head, n_inner=ocfg.n_inner,
                             n_layer=ocfg.n_layer, rotary_emb_base=ocfg.rotary_emb_base)
    model = cb.BiEncoder(cb.BiEncoderConfig(encoder=cfg, hamming=hamming, freeze=freeze, logit_scale=20.0,
                                             trainable_logit_scale=True))
    sd = random_state_dict(ocfg, seed=case["wseed"])  # reference key names (NomicBertModel / its HF copy)
    model.trunk.load_reference_state_dict(sd)
    return model, ocfg, sd


@pytest.mark.parametrize("safe", [True, False])
def test_save_pretrained_layout_and_round_trip(tmp_path, safe):
    model, ocfg, sd = _tiny(hamming=True)
    out = str(tmp_path / "model")
    ck.save_pretrained(model, out, safe_serialization=safe)
    assert sorted(os.listdir(out)) == sorted(["config.json", ck.WEIGHTS_SAFE if safe else ck.WEIGHTS_BIN])
    cfg = json.load(open(os.path.join(out, "config.json")))
    # every constructor field of the reference's BiEncoderConfig (configuration_biencoder.py:5-19) is present
    for k in ("model_name", "projection_dim", "logit_scale", "use_fused_kernels", "pooling", "nomic_encoder", "freeze",
              "trainable_logit_scale", "hamming", "pretrained", "gradient_checkpointing"):
        assert k in cfg, k
    assert cfg["hamming"] is True and cfg["pooling"] == "mean" and cfg["nomic_encoder"] is True
    written = ck.read_state_dict(out)
    # the reference's BiEncoder state dict: the trunk's keys under "trunk." and nothing else for this configuration
    assert sorted(written) == sorted("trunk." + k for k in sd)
    for k, v in sd.items():
        assert torch.equal(written["trunk." + k], v.float()), k
    again = ck.from_pretrained(out)
    assert again.config == model.config
    assert torch.equal(again.trunk._flat, model.trunk._flat)


def test_load_accepts_bare_trunk_keys_and_rejects_foreign_ones(tmp_path):
    model, ocfg, sd = _tiny()
    other, _, _ = _tiny()
    other.trunk._flat.zero_()
    ck.load_weights(other, sd)  # bare NomicBertModel keys (a trunk-only checkpoint such as nomic-bert-2048)
    assert torch.equal(other.trunk._flat, model.trunk._flat)
    with pytest.raises(KeyError):
        ck.load_weights(other, {**{"trunk." + k: v for k, v in sd.items()}, "proj.weight": torch.zeros(2, 2)})
    partial = dict(sd)
    partial.pop("emb_ln.weight")
    with pytest.raises(KeyError):
        ck.load_weights(other, partial)


def test_unsupported_trunk_architectures_fail_loudly(tmp_path):
    model, _, _ = _tiny()
    d = ck.config_to_dict(model.config)
    for key, bad in (("prenorm", True), ("use_rms_norm", True), ("qkv_proj_bias", True), ("rotary_emb_fraction", 0.5),
                     ("activation_function", "gelu")):
        broken = json.loads(json.dumps(d))
        broken["trunk_config"][key] = bad
        with pytest.raises(NotImplementedError):
            ck.config_from_dict(broken)
    no_trunk = {k: v for k, v in d.items() if k != "trunk_config"}
    with pytest.raises(ValueError):
        ck.config_from_dict(no_trunk)


def test_logit_scale_file_only_when_trainable(tmp_path):
    d = str(tmp_path)
    frozen = cb.LogitScale(logit_scale=50.0, trainable_logit_scale=False)
    assert ck.save_logit_scale(frozen, d) is False and not os.path.exists(os.path.join(d, "logit_scale.pt"))
    ls = cb.LogitScale(logit_scale=1 / 0.07, trainable_logit_scale=True)
    assert ck.save_logit_scale(ls, d) is True
    blob = torch.load(os.path.join(d, "logit_scale.pt"), weights_only=True)
    assert list(blob) == ["logit_scale"]  # the reference's key (modeling_biencoder.py:30-41), log space
    assert abs(blob["logit_scale"].exp().item() - 1 / 0.07) < 1e-4
    other = cb.LogitScale(logit_scale=1.0, trainable_logit_scale=True)
    assert ck.load_logit_scale(other, d) is True
    assert torch.equal(other.logit_scale, ls.logit_scale)


def test_save_state_load_state_round_trip(tmp_path):
    model, _, _ = _tiny()
    trunk = model.trunk
    g = torch.Generator().manual_seed(3)
    trunk._opt_state = dict(step=7, m=torch.randn(trunk._n_total, generator=g), v=torch.rand(trunk._n_total, generator=g))
    ls = cb.LogitScale(logit_scale=30.0, trainable_logit_scale=True)
    out = str(tmp_path / "step_7")
    torch.manual_seed(1234)
    np.random.seed(5)
    ck.save_state(out, model, ls, process_index=0, scheduler_state={"last_epoch": 7})
    # the reference's file set (trainers/base.py:316-344)
    assert sorted(os.listdir(out)) == ["model", "optimizer.pt", "random_states_0.pt", "scheduler.pt"]
    assert sorted(os.listdir(os.path.join(out, "model"))) == ["config.json", "logit_scale.pt", "model.safetensors"]
    want_torch, want_np = torch.rand(4), np.random.rand(4)  # what the RNG streams produce right after the save point
    opt = torch.load(os.path.join(out, "optimizer.pt"), weights_only=True)
    # torch.optim.AdamW layout: integer ids in param_groups order (decay names sorted, then no-decay names sorted)
    decay, no_decay = trunk._optimizer_param_order("trunk.")
    assert [g["params"] for g in opt["param_groups"]] == [list(range(len(decay))), list(range(len(decay), len(decay) + len(no_decay)))]
    assert set(opt["state"]) == set(range(len(trunk._offsets))) and float(opt["state"][0]["step"]) == 7.0
    name = "encoder.layers.1.mlp.fc2.weight"
    assert torch.equal(opt["state"][decay.index("trunk." + name)]["exp_avg"], trunk.view(trunk._opt_state["m"], name))

    fresh, _, _ = _tiny()
    fresh.trunk._flat.zero_()
    ls2 = cb.LogitScale(logit_scale=1.0, trainable_logit_scale=True)
    torch.manual_seed(0)
    np.random.seed(0)
    sched = ck.load_state(out, fresh, ls2, process_index=0)
    assert sched == {"last_epoch": 7}
    assert torch.equal(fresh.trunk._flat, trunk._flat)
    assert fresh.trunk._opt_state["step"] == 7
    for name in trunk._offsets:  # padding between tensors is not part of the state
        assert torch.equal(fresh.trunk.view(fresh.trunk._opt_state["m"], name), trunk.view(trunk._opt_state["m"], name))
        assert torch.equal(fresh.trunk.view(fresh.trunk._opt_state["v"], name), trunk.view(trunk._opt_state["v"], name))
    assert torch.equal(ls2.logit_scale, ls.logit_scale)
    assert torch.equal(torch.rand(4), want_torch) and np.array_equal(np.random.rand(4), want_np)  # RNG streams resume


def test_optimizer_pt_loads_into_torch_adamw_and_back(tmp_path):
    """optimizer.pt is exchangeable with the reference trainer: its ``optimizer.load_state_dict`` (trainers/base.py:300-301) on
    the AdamW that ``configure_optimizer`` (optimizer.py:7-47, restated here) builds accepts our file, continues from our
    moments, and the state dict it writes loads back into the fused optimizer."""
    model, _, _ = _tiny()
    trunk = model.trunk
    g = torch.Generator().manual_seed(5)
    trunk._opt_state = dict(step=3, m=torch.randn(trunk._n_total, generator=g), v=torch.rand(trunk._n_total, generator=g),
                            hyper=dict(lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01))
    ours = trunk.optimizer_state_dict(prefix="trunk.")
    # configure_optimizer's grouping on named_parameters() of the BiEncoder
    named = dict(model.named_parameters())
    decay = sorted(n for n, p in named.items() if p.squeeze().ndim >= 2 and "bias" not in n)
    no_decay = sorted(n for n in named if n not in decay)
    opt = torch.optim.AdamW([{"params": [named[n] for n in decay], "weight_decay": 0.01, "lr": 2e-4},
                             {"params": [named[n] for n in no_decay], "weight_decay": 0.0, "lr": 2e-4}], betas=(0.9, 0.999), eps=1e-8)
    opt.load_state_dict(ours)
    name = "trunk.encoder.layers.0.attn.Wqkv.weight"
    st = opt.state[named[name]]
    assert float(st["step"]) == 3.0 and torch.equal(st["exp_avg"], trunk.view(trunk._opt_state["m"], name[len("trunk."):]))
    back = opt.state_dict()
    fresh, _, _ = _tiny()
    fresh.trunk.load_optimizer_state_dict(back, prefix="trunk.")
    assert fresh.trunk._opt_state["step"] == 3 and fresh.trunk._opt_state["hyper"]["lr"] == 2e-4
    for n in trunk._offsets:
        assert torch.equal(fresh.trunk.view(fresh.trunk._opt_state["v"], n), trunk.view(trunk._opt_state["v"], n))


def test_torch_optimizer_two_steps_and_zero_grad_set_to_none():
    """after ``optimizer.zero_grad(set_to_none=True)`` the gradient views were gone and later steps skipped
    every parameter.  The views are re-bound by ``_ensure_grad_views`` (called at the start of each backward) and the flat
    buffer is zeroed when they were found dropped.  CPU check of the plumbing (the kernels are not involved)."""
    model, _, _ = _tiny()
    trunk = model.trunk
    opt = torch.optim.SGD(trunk.parameters(), lr=0.5)
    w = trunk.view(trunk._flat, "encoder.layers.0.attn.Wqkv.weight")
    for step in range(2):
        trunk._ensure_grad_views()                      # what _TrunkFn.backward does first
        trunk._flat_grad.add_(1.0)                      # stand-in for the kernels' accumulation
        before = w.clone()
        opt.step()
        assert torch.allclose(w, before - 0.5), step    # the step saw exactly this step's gradient (not a stale sum)
        opt.zero_grad(set_to_none=True)
        assert next(trunk.parameters()).grad is None
    trunk._ensure_grad_views()
    assert float(trunk._flat_grad.abs().sum()) == 0.0 and next(trunk.parameters()).grad is not None


def test_load_state_dict_invalidates_the_bf16_shadow():
    """nn.Module.load_state_dict writes the master through the views; the shadow version must move."""
    model, _, _ = _tiny()
    v0 = model.trunk._master_version
    model.load_state_dict({k: torch.zeros_like(v) for k, v in model.state_dict().items()})
    assert model.trunk._master_version > v0 and float(model.trunk._flat.abs().sum()) == 0.0


def test_apply_keeps_adam_moments():
    model, _, _ = _tiny()
    trunk = model.trunk
    trunk._opt_state = dict(step=2, m=torch.ones(trunk._n_total), v=torch.ones(trunk._n_total))
    model.to(torch.device("cpu"))
    model.double()  # any _apply: the master stays fp32 and the moments survive
    assert trunk._opt_state is not None and trunk._opt_state["step"] == 2 and trunk._opt_state["m"].dtype == torch.float32


@pytest.mark.skipif(not ref_loader.available(), reason="the reference tree exists in the build container only")
def test_files_are_readable_by_the_reference_classes(tmp_path):
    """config.json parses with the reference's own BiEncoderConfig, and the weights load (strict) into the reference's
    pure-PyTorch NomicBertModel built from the embedded trunk_config."""
    import importlib
    ref = ref_loader.load()
    ref_loader._pkg("contrastors.models.biencoder", os.path.join(ref_loader.REF_ROOT, "models", "biencoder"))
    ref_cfg_mod = importlib.import_module("contrastors.models.biencoder.configuration_biencoder")
    model, ocfg, sd = _tiny(hamming=True)
    out = str(tmp_path / "model")
    ck.save_pretrained(model, out)
    rc = ref_cfg_mod.BiEncoderConfig.from_pretrained(out)
    assert rc.pooling == "mean" and rc.hamming is True and rc.nomic_encoder is True and rc.projection_dim is None
    assert abs(rc.logit_scale - 20.0) < 1e-12 and rc.trainable_logit_scale is True
    tc = dict(rc.trunk_config)
    tc.pop("model_type")
    hf_cfg = ref.hf_cfg.NomicBertConfig(**tc)
    hf_model = ref.hf.NomicBertModel(hf_cfg, add_pooling_layer=False)
    written = {k[len("trunk."):]: v for k, v in ck.read_state_dict(out).items()}
    missing, unexpected = hf_model.load_state_dict(written, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
