"""Flat f32 parameter / Adam-moment storage shared between torch modules and the CUDA kernels.

The kernels (csrc/mlp.cu) address all 11k parameters of the actor-critic through ONE flat buffer
plus the offsets of ``ts_actor_critic_desc``.  To keep ``state_dict()`` / ``load_state_dict()`` /
the Collector's torch forward working unchanged (checkpoint/resume), every
``nn.Parameter`` is re-pointed at a view of that buffer: kernels update the flat buffer in place
and the modules see the new weights without copies.
"""
from __future__ import annotations

import math
from typing import Any

import torch
from torch import nn
from torch.distributions import Independent, Normal

from .._cabi import ActorCriticDesc


class UnsupportedModelError(NotImplementedError):
    """Raised when actor / critic / dist_fn / optimizer are outside the fused kernel family.
    There is deliberately no eager-PyTorch fallback."""


def _single_linear(mlp: Any, what: str) -> nn.Linear:
    mods = list(mlp.model)
    if len(mods) != 1 or not isinstance(mods[0], nn.Linear):
        raise UnsupportedModelError(f"{what}: expected a single Linear head, got {mods}")
    return mods[0]


def _trunk(net: Any, what: str) -> tuple[nn.Linear, nn.Linear]:
    """Net(obs -> 64 -> 64, Tanh) -> its two Linear layers."""
    mlp = getattr(net, "model", None)
    seq = getattr(mlp, "model", None)
    if seq is None:
        raise UnsupportedModelError(f"{what}: preprocess net must be a tianshou_b200 Net/MLP")
    mods = list(seq)
    lin = [m for m in mods if isinstance(m, nn.Linear)]
    act = [m for m in mods if not isinstance(m, nn.Linear)]
    if len(lin) != 2 or len(mods) != 4 or not all(isinstance(a, nn.Tanh) for a in act):
        raise UnsupportedModelError(
            f"{what}: fused kernels support exactly Linear-Tanh-Linear-Tanh trunks, got {mods}")
    if getattr(net, "softmax", False):
        raise UnsupportedModelError(f"{what}: softmax trunk output unsupported")
    return lin[0], lin[1]


def describe_actor_critic(actor: Any, critic: Any) -> tuple[ActorCriticDesc, list[nn.Parameter]]:
    """Validate the module structure and return (desc, parameters in flat-buffer order)."""
    if getattr(actor, "_c_sigma", False) or not hasattr(actor, "sigma_param"):
        raise UnsupportedModelError("actor: conditioned sigma unsupported (need state-independent sigma_param)")
    if not getattr(actor, "_unbounded", False):
        raise UnsupportedModelError("actor: only unbounded=True (mu without tanh) is supported")
    a1, a2 = _trunk(actor.preprocess, "actor")
    a3 = _single_linear(actor.mu, "actor.mu")
    c1, c2 = _trunk(critic.preprocess, "critic")
    c3 = _single_linear(critic.last, "critic.last")
    if getattr(critic, "apply_preprocess_net_to_obs_only", False):
        raise UnsupportedModelError("critic: apply_preprocess_net_to_obs_only unsupported")
    H, obs = a1.out_features, a1.in_features
    act = a3.out_features
    ok = (H == 64 and a2.in_features == H and a2.out_features == H and a3.in_features == H and
          c1.in_features == obs and c1.out_features == H and c2.in_features == H and c2.out_features == H and
          c3.in_features == H and c3.out_features == 1)
    if not ok:
        raise UnsupportedModelError("fused kernels support obs -> 64 -> 64 -> {act, 1} shapes only")
    if obs > 64 or act > 16:
        raise UnsupportedModelError(f"obs_dim {obs} > 64 or act_dim {act} > 16 unsupported")
    for lin in (a1, a2, a3, c1, c2, c3):
        if lin.bias is None:
            raise UnsupportedModelError("Linear layers need a bias")
    params = [a1.weight, a1.bias, a2.weight, a2.bias, a3.weight, a3.bias, actor.sigma_param,
              c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias]
    d = ActorCriticDesc()
    d.obs_dim, d.act_dim, d.hidden = obs, act, H
    off = 0
    for name, p in zip(["a_w1", "a_b1", "a_w2", "a_b2", "a_w3", "a_b3", "a_logstd",
                        "c_w1", "c_b1", "c_w2", "c_b2", "c_w3", "c_b3"], params, strict=True):
        setattr(d, name, off)
        off += p.numel()
    d.n_params = off
    return d, params


def check_gaussian_dist_fn(dist_fn: Any, act_dim: int, device: torch.device) -> None:
    """The kernels hard-wire Independent(Normal(mu, sigma), 1) (mujoco_ppo.py:133-135)."""
    loc = torch.zeros(2, act_dim, device=device)
    d = dist_fn((loc, torch.ones_like(loc)))
    if not (isinstance(d, Independent) and isinstance(d.base_dist, Normal) and d.reinterpreted_batch_ndims == 1):
        raise UnsupportedModelError(f"dist_fn must build Independent(Normal(loc, scale), 1); got {d}")


class FlatParams:
    """Owns the flat parameter, gradient and Adam-moment buffers."""

    def __init__(self, params: list[nn.Parameter], device: torch.device, grad_extra: int) -> None:
        self.params = params
        self.device = device
        self.n = sum(p.numel() for p in params)
        self.flat = torch.empty(self.n, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.n + grad_extra, dtype=torch.float32, device=device)
        # per-CTA partial gradient rows written by ts_ppo_grad (folded by ts_clip_adam_step)
        from .._cabi import load_library
        self.partial_rows = int(load_library().ts_ppo_partial_rows())
        self.partials = torch.zeros((self.partial_rows, self.n + grad_extra), dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.step = torch.zeros(1, dtype=torch.int64, device=device)
        self.weight_image: torch.Tensor | None = None   # set by the algorithm (needs the network descriptor)
        self._ptrs: list[int] = []
        self.adopt()

    def _views(self, buf: torch.Tensor) -> list[torch.Tensor]:
        out, off = [], 0
        for p in self.params:
            out.append(buf[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return out

    def adopt(self) -> None:
        """Copy current parameter values into the flat buffer and re-point ``p.data`` at views."""
        with torch.no_grad():
            for p, v in zip(self.params, self._views(self.flat), strict=True):
                if p.data.data_ptr() != v.data_ptr():
                    v.copy_(p.data.to(self.device, torch.float32))
                    p.data = v
        self._ptrs = [p.data.data_ptr() for p in self.params]

    def ensure_adopted(self) -> None:
        if [p.data.data_ptr() for p in self.params] != self._ptrs:
            self.adopt()

    # -- torch.optim.Adam state interop ----------------------------------------------------
    def export_state(self, optimizer: torch.optim.Optimizer) -> None:
        """Expose the flat moments as the torch optimizer's per-parameter state (views)."""
        step = float(self.step.item())
        if step == 0 and len(optimizer.state) == 0:
            return
        for p, m, v in zip(self.params, self._views(self.exp_avg), self._views(self.exp_avg_sq), strict=True):
            optimizer.state[p] = {"step": torch.tensor(step, dtype=torch.float32), "exp_avg": m, "exp_avg_sq": v}

    def import_state(self, optimizer: torch.optim.Optimizer) -> None:
        """After ``optimizer.load_state_dict``: pull its moments / step into the flat buffers."""
        steps = []
        with torch.no_grad():
            for p, m, v in zip(self.params, self._views(self.exp_avg), self._views(self.exp_avg_sq), strict=True):
                st = optimizer.state.get(p)
                if not st:
                    m.zero_(); v.zero_()
                    continue
                m.copy_(st["exp_avg"].to(self.device, torch.float32))
                v.copy_(st["exp_avg_sq"].to(self.device, torch.float32))
                steps.append(float(st["step"]))
        if steps:
            if max(steps) != min(steps):
                raise UnsupportedModelError("per-parameter Adam step counts differ; cannot fuse")
            self.step.fill_(int(round(steps[0])))
        else:
            self.step.zero_()
        self.export_state(optimizer)


def adam_hyperparams(optimizer: torch.optim.Optimizer) -> dict[str, float]:
    if type(optimizer) is not torch.optim.Adam:
        raise UnsupportedModelError(f"fused update supports torch.optim.Adam only, got {type(optimizer).__name__}")
    if len(optimizer.param_groups) != 1:
        raise UnsupportedModelError("fused update supports a single param group")
    g = optimizer.param_groups[0]
    if g.get("amsgrad") or g.get("maximize") or g.get("decoupled_weight_decay"):
        raise UnsupportedModelError("amsgrad / maximize / decoupled weight decay unsupported")
    lr = g["lr"]
    lr = float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr)
    b1, b2 = g["betas"]
    if not all(math.isfinite(x) for x in (lr, b1, b2)):
        raise ValueError("non-finite Adam hyper-parameters")
    return dict(lr=lr, beta1=float(b1), beta2=float(b2), adam_eps=float(g["eps"]),
                weight_decay=float(g["weight_decay"]))
