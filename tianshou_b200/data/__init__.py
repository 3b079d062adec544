from .batch import Batch
from .buffer import (
    MalformedBufferError,
    PrioritizedReplayBuffer,
    PrioritizedVectorReplayBuffer,
    ReplayBuffer,
    ReplayBufferManager,
    VectorReplayBuffer,
)
from .stats import SequenceSummaryStats
from .utils.converter import to_numpy, to_torch, to_torch_as
from .utils.segtree import SegmentTree

__all__ = [
    "Batch",
    "MalformedBufferError",
    "PrioritizedReplayBuffer",
    "PrioritizedVectorReplayBuffer",
    "ReplayBuffer",
    "ReplayBufferManager",
    "VectorReplayBuffer",
    "SequenceSummaryStats",
    "SegmentTree",
    "to_numpy",
    "to_torch",
    "to_torch_as",
]
