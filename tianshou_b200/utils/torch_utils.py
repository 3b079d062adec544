"""Context managers the trainer uses around collection / update (tianshou/utils/torch_utils.py:13-49)."""
from __future__ import annotations

from collections.abc import Iterator
from contextlib import contextmanager
from typing import Any

import torch
from torch import nn


@contextmanager
def torch_train_mode(module: nn.Module, enabled: bool = True) -> Iterator[None]:
    original = module.training
    try:
        module.train(enabled)
        yield
    finally:
        module.train(original)


@contextmanager
def policy_within_training_step(policy: Any, enabled: bool = True) -> Iterator[None]:
    original = policy.is_within_training_step
    try:
        policy.is_within_training_step = enabled
        yield
    finally:
        policy.is_within_training_step = original


def torch_device(module: nn.Module) -> torch.device:
    try:
        return next(module.parameters()).device
    except StopIteration:
        return torch.device("cpu")
