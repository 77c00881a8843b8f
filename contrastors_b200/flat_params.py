"""Flat-buffer parameter storage shared by the towers (NomicBertModel, ViTModel).

B200-first layout: every parameter is a view into ONE flat fp32 master buffer (>= 2-D weights first = the AdamW decay
group of the reference's ``configure_optimizer``, optimizer.py:7-47; 1-D parameters last), with a flat fp32 gradient
buffer the weight-gradient GEMMs accumulate into directly (TMA reduce-add), a flat bf16 shadow the forward / backward
GEMMs read, and (lazily) two flat Adam moment buffers.  DDP's bucket machinery collapses to collectives over one
buffer; clip + AdamW + zero_grad + the bf16 refresh are two launches.

``nn.Parameter`` objects stay ordinary parameters (reference key names), so ``state_dict()`` / ``load_state_dict()`` /
torch optimizers keep working:
  * a torch optimizer step rewrites the master through the views -> a global optimizer post-step hook re-casts the shadow;
  * ``load_state_dict`` copies into the views -> a post hook re-casts the shadow;
  * ``zero_grad(set_to_none=True)`` (torch's default, the reference trainer's call at trainers/base.py:385) drops
    ``p.grad`` -> every backward re-binds the views first and zeroes the flat buffer when it finds them dropped.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

# Any torch optimizer step may have rewritten the fp32 master weights through the parameter views; a global post-step
# hook bumps this counter so the bf16 shadow is re-cast before the next forward (our fused AdamW refreshes it itself).
_OPT_STEPS = [0]


def _on_any_optimizer_step(optimizer, args, kwargs):
    _OPT_STEPS[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(_on_any_optimizer_step)
except Exception:  # pragma: no cover - very old torch
    pass


def _attach(root: nn.Module, dotted: str, param: nn.Parameter):
    parts = dotted.split(".")
    mod = root
    for name in parts[:-1]:
        if name not in mod._modules:
            mod.add_module(name, nn.Module())
        mod = mod._modules[name]
    mod.register_parameter(parts[-1], param)


class FlatParamModule(nn.Module):
    """Subclasses call ``_init_flat(decay_specs, no_decay_specs)`` with ``(dotted_name, shape)`` lists."""

    def _init_flat(self, decay_specs: List[Tuple[str, tuple]], no_decay_specs: List[Tuple[str, tuple]]):
        self._specs = list(decay_specs) + list(no_decay_specs)
        self._offsets: Dict[str, Tuple[int, int, tuple]] = {}
        off = 0
        for name, shape in self._specs:
            n = math.prod(shape)
            self._offsets[name] = (off, n, tuple(shape))
            off += (n + 63) // 64 * 64  # keep every tensor 256-byte aligned (TMA bases)
        self._n_decay = self._offsets[no_decay_specs[0][0]][0]
        self._n_total = off
        self._flat = torch.zeros(off, dtype=torch.float32)
        self._flat_grad = torch.zeros(off, dtype=torch.float32)
        self._shadow = None          # bf16 copy of _flat, refreshed lazily
        self._shadow_version = None
        self._master_version = 0     # bumped by everything in this class that rewrites the master weights
        self._opt_state = None
        self._leaves = {}
        for name, shape in self._specs:
            _attach(self, name, nn.Parameter(torch.empty(0)))
            self._leaves[name] = self._named_leaf(name)
        self._rebind()
        # nn.Module.load_state_dict copies into the parameter views (= the fp32 master): the bf16 shadow must be re-cast
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_weights_updated())

    # ---------------------------------------------------------------- views
    def _named_leaf(self, dotted):
        mod = self
        parts = dotted.split(".")
        for name in parts[:-1]:
            mod = mod._modules[name]
        return mod, parts[-1]

    def _param(self, name) -> nn.Parameter:
        mod, leaf = self._leaves[name]
        return mod._parameters[leaf]

    def _rebind(self):
        for name, (off, n, shape) in self._offsets.items():
            p = self._param(name)
            p.data = self._flat[off:off + n].view(shape)
            p.grad = self._flat_grad[off:off + n].view(shape)

    def _apply(self, fn, recurse=True):
        flat = fn(self._flat)
        grad = fn(self._flat_grad)
        self._flat = flat.float() if flat.dtype != torch.float32 else flat  # master weights stay fp32
        self._flat_grad = grad.float() if grad.dtype != torch.float32 else grad
        self._shadow, self._shadow_version = None, None
        if self._opt_state is not None:  # the Adam moments follow the weights (.to() / .cuda() mid-run keeps them)
            st = self._opt_state
            self._opt_state = dict(st, m=fn(st["m"]).float(), v=fn(st["v"]).float())
        self._on_apply()
        self._rebind()
        return self

    def _on_apply(self):
        """Hook for subclasses holding device-side caches (RoPE tables)."""

    def _ensure_grad_views(self):
        """Called at the start of every backward: a dropped / foreign ``p.grad`` means "gradients were reset" (see the
        module docstring): zero the flat buffer and bind the views again."""
        base = self._flat_grad.data_ptr()
        for name, (off, n, shape) in self._offsets.items():
            g = self._param(name).grad
            if g is None or g.data_ptr() != base + 4 * off:
                break
        else:
            return
        self._flat_grad.zero_()
        for name, (off, n, shape) in self._offsets.items():
            self._param(name).grad = self._flat_grad[off:off + n].view(shape)

    def zero_grad(self, set_to_none: bool = False):
        """One memset of the flat gradient buffer; with ``set_to_none`` the views are dropped too (the next backward
        re-binds them)."""
        self._flat_grad.zero_()
        if set_to_none:
            for name in self._offsets:
                self._param(name).grad = None

    def view(self, buf, name):
        off, n, shape = self._offsets[name]
        return buf[off:off + n].view(shape)

    def flat_grad(self):
        return self._flat_grad

    def mark_weights_updated(self):
        """Call after writing the fp32 master weights by any route other than a torch optimizer step, ``load_state_dict``
        or ``fused_adamw_step`` (e.g. an EMA written through ``p.data``)."""
        self._master_version += 1

    def load_reference_state_dict(self, sd, strict=True):
        """Load a state dict with the reference's key names."""
        missing = [k for k in self._offsets if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys: {missing}")
        with torch.no_grad():
            for name, (off, n, shape) in self._offsets.items():
                if name in sd:
                    self._flat[off:off + n].copy_(sd[name].reshape(-1).to(self._flat.device, torch.float32))
        self.mark_weights_updated()

    def shadow(self):
        """bf16 weights for the GEMMs; re-cast only when the fp32 master may have changed."""
        ver = (self._master_version, _OPT_STEPS[0])
        if self._shadow is None or self._shadow_version != ver:
            if self._shadow is None:
                self._shadow = torch.empty(self._n_total, device=self._flat.device, dtype=torch.bfloat16)
            ops.cast_f32_bf16(self._flat, self._shadow)
            self._shadow_version = ver
        return self._shadow

    # ---------------------------------------------------------------- optimizer tail on the flat buffers
    def fused_adamw_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=None, grad_scale=1.0,
                         extra_sq_norm=None):
        """clip_grad_norm_ + AdamW (decay on >=2-D weights only, optimizer.py:7-47) + zero_grad + bf16 refresh, fused:
        two launches over the flat buffers, no host sync (the clip coefficient stays on the device).  ``grad_scale``
        multiplies the gradient first (1 / world_size after a SUM all-reduce = DDP's average).  ``extra_sq_norm`` (0-dim device
        tensor): squared gradient norm of parameters living outside this buffer (poolers, projection) that share the global
        clip; the 1-element clip coefficient is returned for them (None without clipping)."""
        if self._opt_state is None:
            self._opt_state = dict(step=0, m=torch.zeros_like(self._flat), v=torch.zeros_like(self._flat))
        st = self._opt_state
        st["step"] += 1
        st["hyper"] = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=float(weight_decay))
        coef = None
        if max_grad_norm is not None and max_grad_norm > 0:
            # the clip threshold applies to the SCALED gradient: ||s g|| <= c  <=>  ||g|| <= c / s
            out2 = ops.grad_clip_coef(self._flat_grad, max_grad_norm / grad_scale)
            if extra_sq_norm is None:
                coef = out2[1:]
            else:  # total norm over both parameter sets (torch.nn.utils.clip_grad_norm_'s 1e-6)
                total = torch.sqrt((out2[0] * grad_scale) ** 2 + extra_sq_norm.to(out2.dtype))
                coef = torch.clamp(max_grad_norm / (total + 1e-6), max=1.0).reshape(1)
        if self._shadow is None:
            self._shadow = torch.empty(self._n_total, device=self._flat.device, dtype=torch.bfloat16)
        nd = self._n_decay
        for lo, hi, wd in ((0, nd, weight_decay), (nd, self._n_total, 0.0)):
            ops.adamw_step(self._flat[lo:hi], self._flat_grad[lo:hi], st["m"][lo:hi], st["v"][lo:hi], self._shadow[lo:hi], lr,
                           betas[0], betas[1], eps, wd, st["step"], grad_scale_dev=coef, grad_scale=grad_scale, zero_grad=True)
        self._master_version += 1
        self._shadow_version = (self._master_version, _OPT_STEPS[0])  # the kernel just refreshed the shadow
        return coef

    # ---------------------------------------------------------------- optimizer.pt in torch.optim.AdamW's layout
    def _optimizer_param_order(self, prefix: str):
        """Parameter order of the reference's ``configure_optimizer`` (optimizer.py:7-47) for this tower under module path
        ``prefix``: group 0 = decay names sorted, group 1 = no-decay names sorted; a parameter lands in no-decay when it
        is < 2-D after ``squeeze()`` or has "bias" in its name."""
        decay, no_decay = [], []
        for name, (off, n, shape) in self._offsets.items():
            squeezed = [s for s in shape if s != 1]
            full = prefix + name
            (no_decay if (len(squeezed) < 2 or "bias" in full) else decay).append(full)
        return sorted(decay), sorted(no_decay)

    def optimizer_state_dict(self, prefix: str = ""):
        """State of the fused AdamW as ``torch.optim.AdamW.state_dict()`` of the optimizer the reference builds for this
        tower (``configure_optimizer``; ``optimizer.pt``, trainers/base.py:300-301,323-324): integer parameter ids in
        ``param_groups`` order, ``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter."""
        decay, no_decay = self._optimizer_param_order(prefix)
        st = self._opt_state
        hyper = (st or {}).get("hyper") or dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
        base = dict(lr=hyper["lr"], betas=tuple(hyper["betas"]), eps=hyper["eps"], amsgrad=False, maximize=False, foreach=None,
                    capturable=False, differentiable=False, fused=None)
        groups = [dict(base, weight_decay=hyper["weight_decay"], params=list(range(len(decay)))),
                  dict(base, weight_decay=0.0, params=list(range(len(decay), len(decay) + len(no_decay))))]
        state = {}
        if st is not None and st["step"] > 0:
            for idx, full in enumerate(decay + no_decay):
                name = full[len(prefix):]
                state[idx] = {"step": torch.tensor(float(st["step"])),
                              "exp_avg": self.view(st["m"], name).detach().cpu().clone(),
                              "exp_avg_sq": self.view(st["v"], name).detach().cpu().clone()}
        return {"state": state, "param_groups": groups}

    def load_optimizer_state_dict(self, sd, prefix: str = ""):
        """Accepts the torch layout above (ours or one written by the reference trainer for the same tower)."""
        if not sd.get("state"):
            self._opt_state = None
            return
        decay, no_decay = self._optimizer_param_order(prefix)
        names = decay + no_decay
        ids = [i for g in sd["param_groups"] for i in g["params"]]
        if len(ids) != len(names):
            raise KeyError(f"optimizer state covers {len(ids)} parameters, this tower has {len(names)}")
        m, v = torch.zeros_like(self._flat), torch.zeros_like(self._flat)
        step = 0
        with torch.no_grad():
            for idx, full in zip(ids, names):
                ent = sd["state"][idx]
                name = full[len(prefix):]
                if tuple(ent["exp_avg"].shape) != tuple(self._offsets[name][2]):
                    raise KeyError(f"optimizer state {idx} has shape {tuple(ent['exp_avg'].shape)}, {full} is {self._offsets[name][2]}")
                self.view(m, name).copy_(ent["exp_avg"].to(m.device, torch.float32))
                self.view(v, name).copy_(ent["exp_avg_sq"].to(v.device, torch.float32))
                step = max(step, int(float(ent["step"])))
        g0 = sd["param_groups"][0]
        hyper = dict(lr=float(g0["lr"]), betas=tuple(float(b) for b in g0["betas"]), eps=float(g0["eps"]),
                     weight_decay=float(g0["weight_decay"]))
        self._opt_state = dict(step=step, m=m, v=v, hyper=hyper)
