"""Asynchronous device mirror of the replay buffer: the device copy maintained by
``add()`` is bit-identical to the host arrays, survives wrap-around / partial ``buffer_ids`` / ``reset``,
goes stale on out-of-band mutations, and ``PPO.update`` gives the same result with and without it."""
import numpy as np
import pytest
import torch

from tianshou_b200.data import Batch, ReplayBuffer, VectorReplayBuffer
from tianshou_b200.utils import policy_within_training_step
from ts_testutil import build_ppo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("obs", "act", "rew", "terminated", "truncated", "done", "obs_next")


def _step(rng, E, obs_dim=17, act_dim=6, p_done=0.05):
    return Batch(obs=rng.standard_normal((E, obs_dim)).astype(np.float32), act=rng.standard_normal((E, act_dim)).astype(np.float32),
                 rew=rng.standard_normal(E), terminated=rng.random(E) < p_done, truncated=rng.random(E) < p_done / 2,
                 obs_next=rng.standard_normal((E, obs_dim)).astype(np.float32), info=Batch())


def _assert_mirror_equals_host(buf):
    cols = buf.device_columns()
    assert cols is not None, "mirror is stale"
    torch.cuda.synchronize()
    for k in KEYS:
        host = np.asarray(buf._meta[k])
        dev = cols[k].cpu().numpy()
        assert dev.shape == host.shape and np.array_equal(dev, host.astype(dev.dtype)), k


def test_mirror_tracks_add_with_wraparound_and_partial_ids():
    rng = np.random.default_rng(0)
    E, T = 64, 8
    buf = VectorReplayBuffer(E * T, E, device=DEV, device_mirror=True)
    for t in range(T + 3):                       # wraps every sub-buffer
        buf.add(_step(rng, E))
    _assert_mirror_equals_host(buf)
    ids = np.array([3, 17, 40])                  # a few envs only (Collector with ready_env_ids)
    for _ in range(5):
        b = _step(rng, len(ids))
        buf.add(b, buffer_ids=ids)
    _assert_mirror_equals_host(buf)
    buf.reset(keep_statistics=True)              # bookkeeping rewinds, arrays stay
    assert buf.device_columns() is not None
    buf.add(_step(rng, E))
    _assert_mirror_equals_host(buf)


def test_mirror_goes_stale_on_out_of_band_writes_and_resyncs():
    rng = np.random.default_rng(1)
    E, T = 16, 4
    buf = VectorReplayBuffer(E * T, E, device=DEV, device_mirror=True)
    for _ in range(T):
        buf.add(_step(rng, E))
    assert buf.device_columns() is not None
    buf.set_array_at_key(np.ones(E * T), "rew")  # not an add(): the device copy no longer reflects the host
    assert buf.device_columns() is None
    buf.sync_device_mirror()
    _assert_mirror_equals_host(buf)
    buf.add(_step(rng, E))                       # and incremental again afterwards
    _assert_mirror_equals_host(buf)


def test_single_buffer_mirror():
    rng = np.random.default_rng(2)
    buf = ReplayBuffer(20, device=DEV, device_mirror=True)
    for _ in range(27):
        s = _step(rng, 1)
        buf.add(Batch(obs=s.obs[0], act=s.act[0], rew=float(s.rew[0]), terminated=bool(s.terminated[0]),
                      truncated=bool(s.truncated[0]), obs_next=s.obs_next[0], info=Batch()))
    _assert_mirror_equals_host(buf)


def test_update_is_identical_with_and_without_mirror():
    E, T = 64, 32
    out = []
    for mirror in (False, True):
        rng = np.random.default_rng(3)
        buf = VectorReplayBuffer(E * T, E, device=DEV, device_mirror=mirror)
        for _ in range(T):
            buf.add(_step(rng, E, p_done=0.02))
        algo, actor, critic = build_ppo(17, 6, DEV, return_scaling=True, recompute_advantage=True, value_clip=True,
                                        advantage_normalization=False, max_grad_norm=0.5)
        np.random.seed(5)
        with policy_within_training_step(algo.policy):
            stats = algo.update(buffer=buf, batch_size=512, repeat=2)
        if mirror:
            assert buf.device_columns() is not None
        out.append((algo._flat.flat.clone(), stats.loss.mean))
    assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
