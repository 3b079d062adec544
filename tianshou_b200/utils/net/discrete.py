"""Discrete-action actor / critic heads (API of tianshou/utils/net/discrete.py:22-123)."""
from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np
import torch
import torch.nn.functional as F

from .common import MLP, ModuleWithVectorOutput


def dist_fn_categorical_from_logits(logits: torch.Tensor) -> torch.distributions.Categorical:
    """Default distribution function for categorical actors (discrete.py:22-26)."""
    return torch.distributions.Categorical(logits=logits)


class DiscreteActor(ModuleWithVectorOutput):
    """preprocess net -> ``last`` MLP -> action values; probabilities when ``softmax_output``
    (discrete.py:29-92)."""

    def __init__(self, *, preprocess_net: ModuleWithVectorOutput, action_shape: Any,
                 hidden_sizes: Sequence[int] = (), softmax_output: bool = True) -> None:
        output_dim = int(np.prod(action_shape))
        super().__init__(output_dim)
        self.preprocess = preprocess_net
        self.last = MLP(input_dim=preprocess_net.get_output_dim(), output_dim=self.output_dim,
                        hidden_sizes=hidden_sizes)
        self.softmax_output = softmax_output

    def get_preprocess_net(self) -> ModuleWithVectorOutput:
        return self.preprocess

    def forward(self, obs: Any, state: Any = None, info: dict[str, Any] | None = None) -> tuple[torch.Tensor, Any]:
        x, hidden = self.preprocess(obs, state)
        x = self.last(x)
        if self.softmax_output:
            x = F.softmax(x, dim=-1)
        return x, hidden


class DiscreteCritic(ModuleWithVectorOutput):
    """V(s): preprocess net -> ``last`` MLP -> ``last_size`` (discrete.py:94-123)."""

    def __init__(self, *, preprocess_net: ModuleWithVectorOutput, hidden_sizes: Sequence[int] = (),
                 last_size: int = 1) -> None:
        super().__init__(output_dim=last_size)
        self.preprocess = preprocess_net
        self.last = MLP(input_dim=preprocess_net.get_output_dim(), output_dim=last_size, hidden_sizes=hidden_sizes)

    def forward(self, obs: Any, state: Any = None, info: dict[str, Any] | None = None) -> torch.Tensor:
        logits, _ = self.preprocess(obs, state=state)
        return self.last(logits)
