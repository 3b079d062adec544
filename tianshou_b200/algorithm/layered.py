"""Actor-critic networks OUTSIDE the fused 17-64-64 kernels' shape envelope, layer by layer on the tensor cores.

``describe_actor_critic`` (flat_params.py) accepts exactly the shapes the persistent tcgen05 / SIMT update kernels were
written for (two-layer 64-wide trunks, obs <= 64).  Everything else that is still a Linear / ReLU | Tanh actor-critic --
wider or deeper trunks, large observations (Humanoid: 376), the reference's shared-trunk discrete PPO net at other widths
-- runs here: every Linear forward / input gradient / weight gradient is one ``ts_net_gemm`` launch (tcgen05,
fp32-faithful), the PPO / A2C loss between them is ``ts_ppo_rows``, the optimiser ``ts_adam_step`` (global-norm clip + Adam).
Same public behaviour as the fused path (ppo.py:146-224, a2c.py:115-153); single GPU.

Reference structures covered: ``ContinuousActorProbabilistic(unbounded=True, conditioned_sigma=False)`` +
``ContinuousCritic`` (utils/net/continuous.py:96-238), ``DiscreteActor(softmax_output=True)`` + ``DiscreteCritic`` on
separate or ONE shared ``Net`` (utils/net/discrete.py:29-123, test/discrete/test_ppo_discrete.py:90-100).
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch
from torch import nn

from .. import ops
from .._cabi import STATS_STRIDE, call, ptr, stream_ptr
from .flat_params import UnsupportedModelError
from .netgraph import ACT_NONE, FlatGroup, FusedStack, compile_sequential, module_layers

_CHUNK = 131072          # rows per forward chunk of the whole-rollout passes (bounds the activation scratch)


class LayeredActorCritic:
    def __init__(self, actor: Any, critic: Any, device: torch.device) -> None:
        self.device = device
        self.categorical = hasattr(actor, "softmax_output")
        if self.categorical and not actor.softmax_output:
            raise UnsupportedModelError("actor: DiscreteActor needs softmax_output=True (Categorical over probabilities)")
        if not self.categorical:
            if getattr(actor, "_c_sigma", False) or not hasattr(actor, "sigma_param"):
                raise UnsupportedModelError("actor: conditioned sigma unsupported (need state-independent sigma_param)")
            if not getattr(actor, "_unbounded", False):
                raise UnsupportedModelError("actor: only unbounded=True (mu without tanh) is supported")
        if getattr(critic, "apply_preprocess_net_to_obs_only", False):
            raise UnsupportedModelError("critic: apply_preprocess_net_to_obs_only unsupported")
        for net, what in ((actor.preprocess, "actor"), (critic.preprocess, "critic")):
            if getattr(net, "softmax", False):
                raise UnsupportedModelError(f"{what}: softmax trunk output unsupported")
        self.shared = actor.preprocess is critic.preprocess
        a_mods = module_layers(actor.preprocess)
        first = next((m for m in a_mods if isinstance(m, nn.Linear)), None)
        if first is None:
            raise UnsupportedModelError("actor trunk has no Linear layer")
        self.obs_dim = int(first.in_features)
        a_trunk = compile_sequential(a_mods, (self.obs_dim,))
        c_trunk = a_trunk if self.shared else compile_sequential(module_layers(critic.preprocess), (self.obs_dim,))
        a_head = compile_sequential(module_layers(actor.last if self.categorical else actor.mu), (a_trunk[-1].out_dim,))
        c_head = compile_sequential(module_layers(critic.last), (c_trunk[-1].out_dim,))
        if a_head[-1].act != ACT_NONE or c_head[-1].act != ACT_NONE or c_head[-1].out_dim != 1:
            raise UnsupportedModelError("heads must end in a linear layer (critic: one output)")
        self.act_dim = int(a_head[-1].out_dim)
        if self.act_dim > 64:
            raise UnsupportedModelError("action width > 64 unsupported")
        seen: set[int] = set()
        params: list[nn.Parameter] = []
        for p in [*actor.parameters(), *critic.parameters()]:      # ActorCritic(actor, critic).parameters() order, shared once
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.group = FlatGroup(params, device)
        covered = {id(L.weight) for L in (*a_trunk, *c_trunk, *a_head, *c_head)} | {id(L.bias) for L in (*a_trunk, *c_trunk, *a_head, *c_head)}
        extra = [p for p in params if id(p) not in covered]
        self.sigma_param = None if self.categorical else actor.sigma_param
        if [id(p) for p in extra] != ([] if self.categorical else [id(self.sigma_param)]):
            raise UnsupportedModelError("actor / critic hold parameters outside the Linear layers")
        self.a_trunk, self.a_head = FusedStack(a_trunk, self.group, "a_trunk"), FusedStack(a_head, self.group, "a_head")
        self.c_trunk = self.a_trunk if self.shared else FusedStack(c_trunk, self.group, "c_trunk")
        self.c_head = FusedStack(c_head, self.group, "c_head")
        self._a_act, self._c_act = a_trunk[-1].act, c_trunk[-1].act
        self._scratch: dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ helpers
    def _buf(self, name: str, shape: tuple[int, ...] | int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        t = self._scratch.get(name)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = self._scratch[name] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _logstd_ptr(self, buf: torch.Tensor) -> int | None:
        return None if self.categorical else buf.data_ptr() + 4 * self.group.offset(self.sigma_param)

    def _act_rows(self, act: torch.Tensor) -> torch.Tensor:
        a = act.reshape(act.shape[0], -1).to(torch.float32)
        return a.reshape(-1).contiguous() if self.categorical else a.contiguous()

    # ------------------------------------------------------------------ whole-rollout passes (no grad)
    def critic_values(self, obs: torch.Tensor, out: torch.Tensor) -> None:
        """out[r] = critic(obs[r])  (a2c.py:123-126) in row chunks."""
        n = obs.shape[0]
        for lo in range(0, n, _CHUNK):
            hi = min(n, lo + _CHUNK)
            t = self.c_trunk.forward(obs[lo:hi], hi - lo, "vp")
            h = self.c_head.forward(t[-1], hi - lo, "vp")
            out[lo:hi].copy_(h[-1].view(-1))

    def actor_logp(self, obs: torch.Tensor, act: torch.Tensor, out: torch.Tensor, hp: Any) -> None:
        """out[r] = log pi(act[r] | obs[r])  (ppo.py:157-161)."""
        n = obs.shape[0]
        a = self._act_rows(act)
        for lo in range(0, n, _CHUNK):
            hi = min(n, lo + _CHUNK)
            t = self.a_trunk.forward(obs[lo:hi], hi - lo, "lp")
            h = self.a_head.forward(t[-1], hi - lo, "lp")
            call("ts_ppo_rows", ptr(h[-1]), None, self._logstd_ptr(self.group.flat), ptr(a[lo:hi]), None, None, None, None, hi - lo,
                 self.act_dim, int(self.categorical), C.byref(hp), hi - lo, None, ptr(out[lo:hi]), None, None, None, None,
                 stream_ptr(self.device))

    def actor_head(self, obs: torch.Tensor) -> torch.Tensor:
        """mu / logits for ``obs`` (Collector-side inference)."""
        t = self.a_trunk.forward(obs, obs.shape[0], "inf")
        return self.a_head.forward(t[-1], obs.shape[0], "inf")[-1]

    # ------------------------------------------------------------------ one optimiser step
    def minibatch_step(self, batch: Any, idx: torch.Tensor, hp: Any, adv_moments: torch.Tensor | None, optimizer: torch.optim.Optimizer,
                       max_grad_norm: float | None, stats_row: torch.Tensor) -> None:
        """Gather the minibatch rows, forward, loss rows, backward, clip + Adam, stats  (ppo.py:179-216, algorithm_base.py:496-500)."""
        st = stream_ptr(self.device)
        B, A = int(idx.numel()), self.act_dim
        obs = ops.gather_rows(batch.obs, idx)
        act = ops.gather_rows(self._act_rows(batch.act), idx)
        adv, ret = ops.gather_rows(batch.adv, idx), ops.gather_rows(batch.returns, idx)
        lpo, vso = ops.gather_rows(batch.logp_old, idx), ops.gather_rows(batch.v_s, idx)
        at = self.a_trunk.forward(obs, B, "up")
        ah = self.a_head.forward(at[-1], B, "up")
        ct = at if self.shared else self.c_trunk.forward(obs, B, "up")
        ch = self.c_head.forward(ct[-1], B, "up")
        logp = self._buf("logp", B)
        dhead, dval = self._buf("dhead", (B, A)), self._buf("dval", (B, 1))
        dls = None if self.categorical else self._buf("dls", (B, A))
        rows = self._buf("loss_rows", (B, 3))
        call("ts_ppo_rows", ptr(ah[-1]), ptr(ch[-1]), self._logstd_ptr(self.group.flat), ptr(act), ptr(adv), ptr(ret), ptr(lpo), ptr(vso),
             B, A, int(self.categorical), C.byref(hp), B, ptr(adv_moments), ptr(logp), ptr(dhead), ptr(dval), ptr(dls), ptr(rows), st)
        call("ts_ppo_rows_stats", ptr(rows), B, C.byref(hp), ptr(stats_row), st)
        # backward: heads -> d loss / d (trunk pre-activation), then the trunk(s)
        dz_a = self._buf("dz_a", (B, self.a_trunk.layers[-1].out_dim))
        self.a_head.backward(ah, dhead, B, "up", input_grad=True, input_act=(self._a_act, at[-1]) if self._a_act != ACT_NONE else None,
                             dx_out=dz_a)
        if self.shared:
            self.c_head.backward(ch, dval, B, "up", input_grad=True, input_act=(self._c_act, ct[-1]) if self._c_act != ACT_NONE else None,
                                 dx_out=dz_a, dx_accumulate=True)
            self.a_trunk.backward(at, dz_a, B, "up", dy_preact=True)
        else:
            dz_c = self._buf("dz_c", (B, self.c_trunk.layers[-1].out_dim))
            self.c_head.backward(ch, dval, B, "up", input_grad=True, input_act=(self._c_act, ct[-1]) if self._c_act != ACT_NONE else None,
                                 dx_out=dz_c)
            self.a_trunk.backward(at, dz_a, B, "up", dy_preact=True)
            self.c_trunk.backward(ct, dz_c, B, "up", dy_preact=True)
        if dls is not None:
            call("ts_net_colsum", ptr(dls), A, B, A, self._logstd_ptr(self.group.grad), 0, st)
        self.group.adam_step(optimizer, max_grad_norm)


def try_layered(actor: Any, critic: Any) -> LayeredActorCritic:
    plist = list(actor.parameters())
    if not plist or plist[0].device.type != "cuda":
        raise UnsupportedModelError("actor/critic must live on a CUDA device; tianshou_b200 has no CPU path")
    return LayeredActorCritic(actor, critic, plist[0].device)


def layered_update(algo: Any, batch: Any, batch_size: int | None, repeat: int) -> torch.Tensor:
    """The repeat x minibatch loop (ppo.py:164-224) on a layered actor-critic; returns the device loss table [steps, 8]."""
    from ..data.batch import NumpyGlobalPermutationJob, minibatch_bounds
    L: LayeredActorCritic = algo._layered
    dev = L.device
    N = batch.obs.shape[0]
    bounds = minibatch_bounds(N, batch_size or N, merge_last=True)
    n_mb = len(bounds)
    hp = algo._loss_hparams()
    stats = torch.zeros((repeat * n_mb, STATS_STRIDE), dtype=torch.float32, device=dev)
    st = stream_ptr(dev)
    if algo.minibatch_shuffle == "device":
        perms = ops.make_permutation(algo._shuffle_seed, algo._shuffle_epoch, repeat, N, dev)
        algo._shuffle_epoch += repeat
        job = None
    else:
        job = getattr(algo, "_perm_job", None)
        own_job = job is None or job.shape != (repeat, N)
        if own_job:
            job = NumpyGlobalPermutationJob(algo._host_perm_rows(repeat, N), repeat)
    try:
        for r in range(repeat):
            if algo.recompute_adv and r > 0:
                algo._add_returns_and_advantages(batch, None, None)
            perm = perms[r] if job is None else job.wait(r).to(dev, non_blocking=True)
            for m, (lo, hi) in enumerate(bounds):
                adv_mom = None
                if algo.advantage_normalization:
                    sums = L._buf("adv_sums", 2, torch.float64)
                    sums.zero_()
                    call("ts_minibatch_adv_sums", ptr(batch.adv), ptr(perm), lo, hi, ptr(sums), st)
                    adv_mom = L._buf("adv_mom", 2)
                    call("ts_adv_moments_finalize", ptr(sums), hi - lo, ptr(adv_mom), st)
                idx = perm[lo:hi].to(torch.int64)
                L.minibatch_step(batch, idx, hp, adv_mom, algo.optim._optim, algo.optim._max_grad_norm, stats[r * n_mb + m])
    finally:
        if job is not None and algo.minibatch_shuffle != "device" and getattr(algo, "_perm_job", None) is not job:
            job.__exit__(None, None, None)
    return stats
