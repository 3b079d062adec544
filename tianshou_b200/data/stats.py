"""Summary statistics of per-minibatch losses (API of tianshou/data/stats.py:13-62)."""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass

import numpy as np


@dataclass(kw_only=True)
class SequenceSummaryStats:
    """mean / std / max / min of a sequence (population std, numpy defaults)."""

    mean: float
    std: float
    max: float
    min: float

    @classmethod
    def from_sequence(cls, sequence: Sequence[float | int] | np.ndarray) -> "SequenceSummaryStats":
        if len(sequence) == 0:
            return cls(mean=0.0, std=0.0, max=0.0, min=0.0)
        return cls(
            mean=float(np.mean(sequence)),
            std=float(np.std(sequence)),
            max=float(np.max(sequence)),
            min=float(np.min(sequence)),
        )

    @classmethod
    def from_single_value(cls, value: float | int) -> "SequenceSummaryStats":
        return cls(mean=value, std=0.0, max=value, min=value)
