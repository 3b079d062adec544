"""Drop-in adapter for the UNMODIFIED reference ``Collector`` / ``trainer.py``.

The reference has no plugin interface: its callers hard-code ``isinstance`` gates against their own classes
(``isinstance(policy, Algorithm)`` -> ``.policy``, data/collector.py:360; ``isinstance(buf, ReplayBufferManager)``,
:375,:384).  ``install_into_reference()`` widens exactly those gates -- the names the reference's ``collector`` module
resolves at call time -- to accept this package's classes as well.  Nothing in the reference is edited, subclassed or
re-implemented; everything else goes through duck typing (``buffer.add`` adopts the Collector's ``Batch`` entries, the
trainer only calls ``algorithm.update(...)`` / reads the returned stats dataclass).

    import tianshou                      # the reference
    import tianshou_b200.compat as compat
    compat.install_into_reference()
    collector = tianshou.data.Collector(b200_algorithm, envs, tianshou_b200.data.VectorReplayBuffer(...))
    tianshou.trainer.OnPolicyTrainer(b200_algorithm, params).run()
"""
from __future__ import annotations

from typing import Any


def install_into_reference(reference_collector_module: Any = None) -> None:
    """Widen the reference Collector's isinstance gates to this package's ``Algorithm`` / ``ReplayBufferManager``.
    Idempotent.  ``reference_collector_module`` defaults to ``tianshou.data.collector`` (must be importable)."""
    if reference_collector_module is None:
        import importlib
        reference_collector_module = importlib.import_module("tianshou.data.collector")
    from .algorithm.base import Algorithm
    from .data.buffer.base import ReplayBufferManager

    def widen(name: str, ours: type) -> None:
        cur = getattr(reference_collector_module, name)
        group = cur if isinstance(cur, tuple) else (cur,)
        if ours not in group:
            setattr(reference_collector_module, name, (*group, ours))

    widen("Algorithm", Algorithm)
    widen("ReplayBufferManager", ReplayBufferManager)
