"""Synthetic HalfCheetah-shaped rollouts and the MuJoCo-PPO model of the benchmark configuration
(BASELINE.md section 3 / SURVEY.md 8d; model + init from examples/mujoco/mujoco_ppo.py:90-120).
Used by bench.py, smoke() and the tests; there is no dataset or checkpoint involved."""
from __future__ import annotations

from collections.abc import Iterator
from typing import Any

import numpy as np
import torch

MUJOCO_PPO_KWARGS = dict(  # examples/mujoco/mujoco_ppo.py:28-62
    gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.25, ent_coef=0.0, return_scaling=True,
    eps_clip=0.2, value_clip=True, dual_clip=None, advantage_normalization=False, recompute_advantage=True)


class BoxSpace:
    """Duck-typed stand-in for gymnasium.spaces.Box(-1, 1, (dim,))."""

    def __init__(self, dim: int):
        self.shape = (dim,)
        self.low = -np.ones(dim, np.float32)
        self.high = np.ones(dim, np.float32)


def synth_rollout(rng: np.random.Generator, E: int, steps: int, obs_dim: int, act_dim: int,
                  p_term: float = 1e-3, trunc_len: int = 1000) -> Iterator[dict[str, np.ndarray]]:
    """Per-step dicts of [E, ...] arrays: obs/obs_next ~ N(0,1) f32 (obs_next of t is obs of t+1,
    fresh draw after done), act ~ N(0,1) f32, rew ~ N(0,1) f64, terminated ~ Bernoulli(p_term),
    truncated when the episode reaches trunc_len."""
    t_in_ep = np.zeros(E, dtype=np.int64)
    obs = rng.standard_normal((E, obs_dim)).astype(np.float32)
    for _ in range(steps):
        act = rng.standard_normal((E, act_dim)).astype(np.float32)
        rew = rng.standard_normal(E)
        obs_next = rng.standard_normal((E, obs_dim)).astype(np.float32)
        term = rng.random(E) < p_term
        t_in_ep += 1
        trunc = (t_in_ep >= trunc_len) & ~term
        yield dict(obs=obs, act=act, rew=rew, terminated=term, truncated=trunc, obs_next=obs_next)
        done = term | trunc
        t_in_ep[done] = 0
        obs = np.where(done[:, None], rng.standard_normal((E, obs_dim)).astype(np.float32), obs_next)


def gaussian_dist(loc_scale: tuple[torch.Tensor, torch.Tensor]) -> torch.distributions.Distribution:
    loc, scale = loc_scale
    return torch.distributions.Independent(torch.distributions.Normal(loc, scale), 1)


def build_mujoco_actor_critic(obs_dim: int, act_dim: int, device: Any, seed: int = 0):
    from .utils.net.common import Net
    from .utils.net.continuous import ContinuousActorProbabilistic, ContinuousCritic
    torch.manual_seed(seed)
    net_a = Net(state_shape=(obs_dim,), hidden_sizes=(64, 64), activation=torch.nn.Tanh)
    actor = ContinuousActorProbabilistic(preprocess_net=net_a, action_shape=(act_dim,), unbounded=True).to(device)
    net_c = Net(state_shape=(obs_dim,), hidden_sizes=(64, 64), activation=torch.nn.Tanh)
    critic = ContinuousCritic(preprocess_net=net_c).to(device)
    torch.nn.init.constant_(actor.sigma_param, -0.5)
    for m in list(actor.modules()) + list(critic.modules()):
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.orthogonal_(m.weight, gain=np.sqrt(2))
            torch.nn.init.zeros_(m.bias)
    for m in actor.mu.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.zeros_(m.bias)
            m.weight.data.copy_(0.01 * m.weight.data)
    return actor, critic


def build_mujoco_ppo(obs_dim: int, act_dim: int, device: Any, lr: float = 3e-4, seed: int = 0, **overrides: Any):
    from .algorithm import PPO, AdamOptimizerFactory, ProbabilisticActorPolicy
    actor, critic = build_mujoco_actor_critic(obs_dim, act_dim, device, seed)
    policy = ProbabilisticActorPolicy(actor=actor, dist_fn=gaussian_dist, action_scaling=True,
                                      action_bound_method="clip", action_space=BoxSpace(act_dim))
    kw = dict(MUJOCO_PPO_KWARGS)
    kw.update(overrides)
    return PPO(policy=policy, critic=critic, optim=AdamOptimizerFactory(lr=lr), **kw), actor, critic


def fill_vector_buffer(buf: Any, rng: np.random.Generator, E: int, steps: int, obs_dim: int, act_dim: int,
                       **kw: Any) -> None:
    from .data import Batch
    ids = np.arange(E)
    for s in synth_rollout(rng, E, steps, obs_dim, act_dim, **kw):
        buf.add(Batch(**s), buffer_ids=ids)
