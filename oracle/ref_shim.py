"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference (thu-ml/tianshou 2.0.1)
from /root/reference inside the build container.

The reference needs third-party packages that are not installed here (gymnasium, sensai-utils,
overrides, h5py, deepdiff, matplotlib, pettingzoo, rliable).  None of them take part in the
policy-update hot path, so this module registers minimal stand-ins in ``sys.modules`` before
importing ``tianshou``.  Nothing in the product package (``tianshou_b200``) may import this
file; it is used by ``oracle/gen_golden.py`` (fixture generation) and by tests that are skipped
when ``/root/reference`` does not exist (e.g. on the GPU box).

Recipe verified by the survey.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TIANSHOU_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tianshou"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _install_stubs() -> None:
    import numpy as np

    # ---- gymnasium -----------------------------------------------------------------
    class Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self.shape = shape
            self.dtype = dtype

        def seed(self, seed=None):
            return [seed]

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
            super().__init__(tuple(shape), dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            return bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Discrete(Space):
        def __init__(self, n, seed=None, start=0):
            self.n = int(n)
            self.start = start
            super().__init__((), np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec)
            super().__init__(self.nvec.shape, dtype)

    class MultiBinary(Space):
        def __init__(self, n, seed=None):
            self.n = n
            super().__init__((n,), np.int8)

    class Dict(Space, dict):
        pass

    class Tuple(Space, tuple):
        pass

    class Env:
        metadata: dict = {}
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            return getattr(self.env, name)

    class ObservationWrapper(Wrapper):
        pass

    class RewardWrapper(Wrapper):
        pass

    class ActionWrapper(Wrapper):
        pass

    spaces = _mod(
        "gymnasium.spaces",
        Space=Space,
        Box=Box,
        Discrete=Discrete,
        MultiDiscrete=MultiDiscrete,
        MultiBinary=MultiBinary,
        Dict=Dict,
        Tuple=Tuple,
    )
    gym = _mod(
        "gymnasium",
        Env=Env,
        Wrapper=Wrapper,
        ObservationWrapper=ObservationWrapper,
        RewardWrapper=RewardWrapper,
        ActionWrapper=ActionWrapper,
        Space=Space,
        spaces=spaces,
        __version__="0.29.1",
    )
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("gymnasium stub: no envs"))
    sys.modules["gymnasium.spaces"] = spaces
    _mod("gymnasium.spaces.discrete", Discrete=Discrete)
    _mod("gymnasium.spaces.box", Box=Box)
    _mod("gymnasium.core", Env=Env, Wrapper=Wrapper)
    _mod("gymnasium.envs")
    _mod("gymnasium.envs.registration", EnvSpec=type("EnvSpec", (), {}))

    # ---- sensai --------------------------------------------------------------------
    class ToStringMixin:
        def __repr__(self):
            return f"{self.__class__.__name__}[{id(self)}]"

    def setstate(cls, obj, state, new_default_properties=None, **_kw):
        if new_default_properties:
            for k, v in new_default_properties.items():
                state.setdefault(k, v)
        obj.__dict__.update(state)

    _mod("sensai")
    _mod("sensai.util")
    _mod("sensai.util.hash", pickle_hash=lambda o: hash(repr(o)))
    _mod(
        "sensai.util.helper",
        mark_used=lambda *a, **k: None,
        count_none=lambda *a: sum(x is None for x in a),
    )
    _mod("sensai.util.pickle", setstate=setstate, dump_pickle=None, load_pickle=None)
    _mod("sensai.util.string", ToStringMixin=ToStringMixin, dict_string=lambda d: str(d))
    sl = _mod("sensai.util.logging")
    for k in dir(logging):
        if not k.startswith("__"):
            setattr(sl, k, getattr(logging, k))
    sl.set_configure_callback = lambda *a, **k: None
    sl.datetime_tag = lambda: "stub"
    sl.run_cli = lambda fn: fn()
    _mod("sensai.util.git", git_status=lambda *a, **k: None, GitStatus=type("GitStatus", (), {}))

    # ---- small ones ----------------------------------------------------------------
    def override(f=None, **_k):
        return f if f is not None else (lambda g: g)

    _mod("overrides", override=override, overrides=override)

    class _H5:
        class Dataset:  # noqa: D106
            pass

        class Group:  # noqa: D106
            pass

        class File:  # noqa: D106
            def __init__(self, *a, **k):
                raise RuntimeError("h5py stub")

    _mod("h5py", Dataset=_H5.Dataset, Group=_H5.Group, File=_H5.File)

    def DeepDiff(a, b, **_k):  # structural equality on dicts of arrays
        def eq(x, y):
            if isinstance(x, dict) and isinstance(y, dict):
                return x.keys() == y.keys() and all(eq(x[k], y[k]) for k in x)
            try:
                return bool(np.array_equal(np.asarray(x), np.asarray(y)))
            except Exception:
                return x == y

        return {} if eq(a, b) else {"diff": True}

    _mod("deepdiff", DeepDiff=DeepDiff)
    _mod("matplotlib")
    _mod("matplotlib.figure", Figure=type("Figure", (), {}))
    _mod("matplotlib.pyplot")
    _mod("matplotlib.axes", Axes=type("Axes", (), {}))
    _mod("matplotlib.ticker")
    _mod("pettingzoo", __version__="1.24.0")
    _mod("pettingzoo.utils")
    _mod("pettingzoo.utils.env", AECEnv=type("AECEnv", (), {}))
    _mod("pettingzoo.utils.wrappers", BaseWrapper=type("BaseWrapper", (), {}))
    _mod("rliable")
    _mod("rliable.library")
    _mod("rliable.plot_utils")
    sys.modules["sensai.util"].logging = sys.modules["sensai.util.logging"]


_REF = None


def import_reference():
    """Return the imported reference ``tianshou`` package (cached)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("gymnasium", "sensai", "overrides", "h5py", "deepdiff", "pettingzoo"):
        try:
            importlib.import_module(name)
        except Exception:
            _install_stubs()
            break
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _REF = importlib.import_module("tianshou")
    return _REF
