"""to_numpy / to_torch / to_torch_as (API of tianshou/data/utils/converter.py:17-100)."""
from __future__ import annotations

from copy import deepcopy
from numbers import Number
from typing import Any

import numpy as np
import torch

from ..batch import Batch, _coerce


def to_numpy(x: Any) -> Batch | np.ndarray:
    """Recursively turn tensors into numpy arrays; numbers / None become (object) arrays."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, (np.number, np.bool_, Number)):
        return np.asanyarray(x)
    if x is None:
        return np.array(None, dtype=object)
    if isinstance(x, (dict, Batch)):
        out = Batch(x, copy=True) if isinstance(x, dict) else deepcopy(x)
        out.to_numpy_()
        return out
    if isinstance(x, (list, tuple)):
        return to_numpy(_coerce(x))
    return np.asanyarray(x)


def to_torch(
    x: Any,
    dtype: torch.dtype | None = None,
    device: str | int | torch.device = "cpu",
) -> Batch | torch.Tensor:
    """Recursively turn numpy arrays / numbers into tensors on ``device``."""
    if isinstance(x, np.ndarray) and issubclass(x.dtype.type, (np.bool_, np.number)):
        t = torch.from_numpy(x).to(device)
        return t.type(dtype) if dtype is not None else t
    if isinstance(x, torch.Tensor):
        if dtype is not None:
            x = x.type(dtype)
        return x.to(device)
    if isinstance(x, (np.number, np.bool_, Number)):
        return to_torch(np.asanyarray(x), dtype, device)
    if isinstance(x, (dict, Batch)):
        out = Batch(x, copy=True) if isinstance(x, dict) else deepcopy(x)
        out.to_torch_(dtype, device)
        return out
    if isinstance(x, (list, tuple)):
        return to_torch(_coerce(x), dtype, device)
    raise TypeError(f"object {x} cannot be converted to torch.")


def to_torch_as(x: Any, y: torch.Tensor) -> Batch | torch.Tensor:
    """``to_torch(x, dtype=y.dtype, device=y.device)`` (converter.py:95-100)."""
    assert isinstance(y, torch.Tensor)
    return to_torch(x, dtype=y.dtype, device=y.device)
