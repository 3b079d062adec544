from .base import MalformedBufferError, ReplayBuffer, ReplayBufferManager, VectorReplayBuffer
from .prio import PrioritizedReplayBuffer, PrioritizedVectorReplayBuffer

__all__ = [
    "MalformedBufferError",
    "ReplayBuffer",
    "ReplayBufferManager",
    "VectorReplayBuffer",
    "PrioritizedReplayBuffer",
    "PrioritizedVectorReplayBuffer",
]
