"""Optimizer / LR-scheduler factories (API of tianshou/algorithm/optim.py:16-140)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Callable, Iterable
from typing import Any

import numpy as np
import torch
from torch.optim import Adam, RMSprop
from torch.optim.lr_scheduler import LambdaLR, LRScheduler


class LRSchedulerFactory(ABC):
    @abstractmethod
    def create_scheduler(self, optim: torch.optim.Optimizer) -> LRScheduler: ...


class LRSchedulerFactoryLinear(LRSchedulerFactory):
    """lr * (1 - n_updates / max_updates), max_updates = ceil(epoch_steps / collect_steps) * epochs."""

    def __init__(self, max_epochs: int, epoch_num_steps: int, collection_step_num_env_steps: int):
        self.num_epochs = max_epochs
        self.epoch_num_steps = epoch_num_steps
        self.collection_step_num_env_steps = collection_step_num_env_steps

    def create_scheduler(self, optim: torch.optim.Optimizer) -> LRScheduler:
        max_update_num = np.ceil(self.epoch_num_steps / self.collection_step_num_env_steps) * self.num_epochs
        return LambdaLR(optim, lr_lambda=lambda epoch: 1.0 - epoch / max_update_num)


class OptimizerFactory(ABC):
    def __init__(self) -> None:
        self.lr_scheduler_factory: LRSchedulerFactory | None = None

    def with_lr_scheduler_factory(self, lr_scheduler_factory: LRSchedulerFactory) -> "OptimizerFactory":
        self.lr_scheduler_factory = lr_scheduler_factory
        return self

    def create_instances(self, module: torch.nn.Module) -> tuple[torch.optim.Optimizer, LRScheduler | None]:
        optimizer = self._create_optimizer_for_params(module.parameters())
        sched = None
        if self.lr_scheduler_factory is not None:
            sched = self.lr_scheduler_factory.create_scheduler(optimizer)
        return optimizer, sched

    @abstractmethod
    def _create_optimizer_for_params(self, params: Iterable[Any]) -> torch.optim.Optimizer: ...


class TorchOptimizerFactory(OptimizerFactory):
    def __init__(self, optim_class: Callable[..., torch.optim.Optimizer], **kwargs: Any):
        super().__init__()
        self.optim_class = optim_class
        self.kwargs = kwargs

    def _create_optimizer_for_params(self, params: Iterable[Any]) -> torch.optim.Optimizer:
        return self.optim_class(params, **self.kwargs)


class AdamOptimizerFactory(OptimizerFactory):
    def __init__(self, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-08,
                 weight_decay: float = 0):
        super().__init__()
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay

    def _create_optimizer_for_params(self, params: Iterable[Any]) -> torch.optim.Optimizer:
        return Adam(params, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)


class RMSpropOptimizerFactory(OptimizerFactory):
    def __init__(self, lr: float = 1e-2, alpha: float = 0.99, eps: float = 1e-08, weight_decay: float = 0,
                 momentum: float = 0, centered: bool = False):
        super().__init__()
        self.kw = dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum,
                       centered=centered)

    def _create_optimizer_for_params(self, params: Iterable[Any]) -> torch.optim.Optimizer:
        return RMSprop(params, **self.kw)
