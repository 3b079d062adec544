"""``SegmentTree`` whose f64 tree lives in HBM; every operation is a CUDA kernel behind the C ABI.

API of tianshou/data/utils/segtree.py:5-82 (``tree[idx]``, ``tree[idx] = v``, ``reduce``,
``get_prefix_sum_idx``) with numpy in / numpy out so that ``PrioritizedReplayBuffer`` and the
reference's own tests can drive it unchanged.  Device-tensor fast paths (``*_device``) avoid the
host round trip for callers that stay on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from ..._cabi import require_cuda, to_device


class SegmentTree:
    def __init__(self, size: int, device: torch.device | str | None = None) -> None:
        bound = 1
        while bound < size:
            bound *= 2
        self._size = size
        self._bound = bound
        self._device_arg = device
        self._tree: torch.Tensor | None = None  # allocated on first device use

    # -- device state ---------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.tree.device

    @property
    def tree(self) -> torch.Tensor:
        if self._tree is None:
            require_cuda()
            dev = torch.device(self._device_arg) if self._device_arg is not None else torch.device(
                "cuda", torch.cuda.current_device())
            self._tree = torch.zeros(2 * self._bound, dtype=torch.float64, device=dev)
        return self._tree

    @property
    def bound(self) -> int:
        return self._bound

    def __len__(self) -> int:
        return self._size

    # -- reference API (host in / host out) ---------------------------------------------------
    def __getitem__(self, index: int | np.ndarray) -> float | np.ndarray:
        if isinstance(index, (int, np.integer)):
            return float(self.tree[int(index) + self._bound].item())
        idx = to_device(np.asarray(index, dtype=np.int64) + self._bound, self.device)
        return ops.gather_rows(self.tree, idx).cpu().numpy().reshape(np.shape(index))

    def __setitem__(self, index: int | np.ndarray, value: float | np.ndarray) -> None:
        if isinstance(index, (int, np.integer)):
            index, value = np.array([index]), np.array([value])
        index = np.asarray(index, dtype=np.int64).reshape(-1)
        assert np.all(index >= 0)
        assert np.all(index < self._size)
        value = np.asarray(value)
        if value.ndim == 0:
            value = np.full(index.shape, float(value))
        if value.dtype not in (np.float32, np.float64):
            value = value.astype(np.float64)
        self.setitem_device(to_device(index, self.device), to_device(value.reshape(-1), self.device))

    def reduce(self, start: int = 0, end: int | None = None) -> float:
        if start == 0 and end is None:
            return float(self.tree[1].item())
        if end is None:
            end = self._size
        if end < 0:
            end += self._size
        return float(ops.segtree_reduce(self.tree, self._bound, start, end).item())

    def get_prefix_sum_idx(self, value: float | np.ndarray) -> int | np.ndarray:
        single = not isinstance(value, np.ndarray)
        v = np.atleast_1d(np.asarray(value, dtype=np.float64))
        assert np.all(v >= 0.0)
        assert np.all(v < self.reduce())
        out = self.prefix_sum_idx_device(to_device(v.reshape(-1), self.device)).cpu().numpy()
        return int(out[0]) if single else out.reshape(v.shape)

    # -- device fast paths -----------------------------------------------------------------
    def setitem_device(self, index: torch.Tensor, value: torch.Tensor) -> None:
        n, cap = index.numel(), 32 * 1024
        for lo in range(0, n, cap):  # kernel handles <= 32768 items per launch
            ops.segtree_setitem(self.tree, self._bound, index[lo:lo + cap].contiguous(),
                                value[lo:lo + cap].contiguous())

    def prefix_sum_idx_device(self, value: torch.Tensor) -> torch.Tensor:
        return ops.segtree_prefix_sum_idx(self.tree, self._bound, value)

    def sample_device(self, u: torch.Tensor) -> torch.Tensor:
        """indices for ``u * total`` (prio.py:65-66), u in [0,1) f64 on device."""
        return ops.segtree_sample(self.tree, self._bound, u)
