"""Stochastic actor policy (API of tianshou/algorithm/modelfree/reinforce.py:62-192)."""
from __future__ import annotations

import warnings
from collections.abc import Callable
from typing import Any, Literal

import numpy as np
import torch

from ...data import Batch
from ..base import Policy

TDistFn = Callable[..., torch.distributions.Distribution]


class ProbabilisticActorPolicy(Policy):
    """actor(obs) -> dist_fn(...) -> sampled (or modal) action."""

    def __init__(self, *, actor: torch.nn.Module, dist_fn: TDistFn, deterministic_eval: bool = False,
                 action_space: Any, observation_space: Any | None = None, action_scaling: bool = True,
                 action_bound_method: Literal["clip", "tanh"] | None = "clip") -> None:
        super().__init__(action_space=action_space, observation_space=observation_space,
                         action_scaling=action_scaling, action_bound_method=action_bound_method)
        if action_scaling:
            try:
                if np.isclose(float(actor.max_action), 1.0) and not getattr(actor, "_unbounded", False):
                    warnings.warn(
                        "action_scaling and action_bound_method are only intended to deal with unbounded "
                        "model action space; consider unbounded=True for the actor.")
            except BaseException:
                pass
        self.actor = actor
        self.dist_fn = dist_fn
        self._eps = 1e-8
        self.deterministic_eval = deterministic_eval

    def forward(self, batch: Batch, state: Any = None) -> Batch:
        """Batch(logits, act, state, dist); samples unless deterministic evaluation applies
        (reinforce.py:167-192)."""
        dist_input, hidden = self.actor(batch.obs, state=state, info=batch.get("info"))
        dist = self.dist_fn(dist_input)
        if self.deterministic_eval and not self.is_within_training_step:
            act = dist.mode
        else:
            act = dist.sample()
        return Batch(logits=dist_input, act=act, state=hidden, dist=dist)


class DiscreteActorPolicy(ProbabilisticActorPolicy):
    """Categorical policy over a Discrete action space (reinforce.py:195-243): no action scaling / bounding."""

    def __init__(self, *, actor: torch.nn.Module, dist_fn: TDistFn | None = None, deterministic_eval: bool = False,
                 action_space: Any, observation_space: Any | None = None) -> None:
        if not (type(action_space).__name__ == "Discrete" or hasattr(action_space, "n")):
            raise ValueError(f"Action space must be an instance of Discrete; got {action_space}")
        if dist_fn is None:
            from ...utils.net.discrete import dist_fn_categorical_from_logits
            dist_fn = dist_fn_categorical_from_logits
        super().__init__(actor=actor, dist_fn=dist_fn, deterministic_eval=deterministic_eval, action_space=action_space,
                         observation_space=observation_space, action_scaling=False, action_bound_method=None)
