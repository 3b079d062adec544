"""Prioritized replay on top of the device-resident sum tree.

Contract: tianshou/data/buffer/prio.py:12-113 and manager.py:237-256 / vecbuf.py:40-66.
The uniform scalars stay a host draw from the *global* numpy RNG (``np.random.rand``, prio.py:65)
so index streams are identical to the reference; the tree, the prefix-sum descent, the
importance weights and the priority updates are CUDA kernels (csrc/segtree.cu).
``weight ** alpha`` is evaluated with numpy on the <= batch-size host array before upload so the
tree holds the same f64 bits as the reference's (glibc ``pow``), keeping sampled indices
bit-exact.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from ..batch import Batch, IndexType
from ..utils.converter import to_numpy
from ..utils.segtree import SegmentTree
from .base import ReplayBuffer, VectorReplayBuffer


class _PrioritizedMixin:
    def _init_prio(self, size: int, alpha: float, beta: float, weight_norm: bool) -> None:
        assert alpha > 0.0
        assert beta >= 0.0
        d = self.__dict__
        d["_alpha"], d["_beta"] = alpha, beta
        d["_max_prio"] = d["_min_prio"] = 1.0
        d["weight"] = SegmentTree(size, device=d.get("_device_arg"))
        d["_prio_eps"] = np.finfo(np.float32).eps.item()
        d["_weight_norm"] = weight_norm
        self.options.update(alpha=alpha, beta=beta)  # type: ignore[attr-defined]

    def init_weight(self, index: int | np.ndarray) -> None:
        self.weight[index] = self._max_prio ** self._alpha  # type: ignore[attr-defined]

    def update(self, buffer: ReplayBuffer) -> np.ndarray:
        indices = super().update(buffer)  # type: ignore[misc]
        self.init_weight(indices)
        return indices

    def add(self, batch: Batch, buffer_ids: np.ndarray | list[int] | None = None):
        ptr, ep_rew, ep_len, ep_idx = super().add(batch, buffer_ids)  # type: ignore[misc]
        self.init_weight(ptr)
        return ptr, ep_rew, ep_len, ep_idx

    def sample_indices(self, batch_size: int | None) -> np.ndarray:
        if batch_size is not None and batch_size > 0 and len(self) > 0:  # type: ignore[arg-type]
            scalar = np.random.rand(batch_size) * self.weight.reduce()  # type: ignore[attr-defined]
            return self.weight.get_prefix_sum_idx(scalar)  # type: ignore[attr-defined]
        return super().sample_indices(batch_size)  # type: ignore[misc]

    def get_weight(self, index: int | np.ndarray) -> float | np.ndarray:
        """(p_i / min_prio) ** (-beta)  -- note min_prio is not alpha-exponentiated (prio.py:69-79)."""
        return (self.weight[index] / self._min_prio) ** (-self._beta)  # type: ignore[attr-defined]

    def update_weight(self, index: np.ndarray, new_weight: np.ndarray | torch.Tensor) -> None:
        weight = np.abs(to_numpy(new_weight)) + self._prio_eps  # type: ignore[attr-defined]
        self.weight[index] = weight ** self._alpha  # type: ignore[attr-defined]
        self.__dict__["_max_prio"] = max(self._max_prio, weight.max())  # type: ignore[attr-defined]
        self.__dict__["_min_prio"] = min(self._min_prio, weight.min())  # type: ignore[attr-defined]

    def __getitem__(self, index: IndexType) -> Batch:
        if isinstance(index, slice):
            indices = self.sample_indices(0) if index == slice(None) else self._indices[: len(self)][index]  # type: ignore
        else:
            indices = index
        batch = super().__getitem__(indices)  # type: ignore[misc]
        weight = self.get_weight(indices)
        batch.weight = weight / np.max(weight) if self._weight_norm else weight  # type: ignore[attr-defined]
        return batch

    def set_beta(self, beta: float) -> None:
        self.__dict__["_beta"] = beta


class PrioritizedReplayBuffer(_PrioritizedMixin, ReplayBuffer):
    def __init__(self, size: int, alpha: float, beta: float, weight_norm: bool = True, **kwargs: Any) -> None:
        ReplayBuffer.__init__(self, size, **kwargs)
        self._init_prio(size, alpha, beta, weight_norm)


class PrioritizedVectorReplayBuffer(_PrioritizedMixin, VectorReplayBuffer):
    """One sum tree over all ``buffer_num`` sub-buffers (manager.py:237-256, vecbuf.py:40-66)."""

    def __init__(self, total_size: int, buffer_num: int, alpha: float, beta: float,
                 weight_norm: bool = True, **kwargs: Any) -> None:
        VectorReplayBuffer.__init__(self, total_size, buffer_num, **kwargs)
        self._init_prio(self.maxsize, alpha, beta, weight_norm)
