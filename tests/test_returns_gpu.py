"""The reference's own return tests (test/base/test_returns.py) driven through tianshou_b200's
ReplayBuffer / Algorithm API: same construction, same known answers."""
import numpy as np
import pytest
import torch

from test_oracle import GAE_KATS, NSTEP_KATS
from ts_testutil import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("k", range(len(GAE_KATS)))
def test_episodic_returns(k):
    from tianshou_b200.algorithm import Algorithm
    from tianshou_b200.data import Batch, ReplayBuffer
    term, trunc, rew, v, gamma, lam, ans = GAE_KATS[k]
    buf = ReplayBuffer(20, device=DEV)
    batch = Batch(terminated=np.array(term, dtype=float), truncated=np.array(trunc, dtype=float),
                  rew=np.array(rew, dtype=float))
    for b in batch:
        b.obs = b.act = 1
        buf.add(b)
    v_t = None if v is None else torch.tensor(v, dtype=torch.float64)
    returns, _ = Algorithm.compute_episodic_return(batch, buf, buf.sample_indices(0), v_t, gamma=gamma, gae_lambda=lam)
    assert np.allclose(returns, ans)


def target_q_fn(buffer, indices):
    indices = buffer.next(indices)
    return torch.tensor(-buffer.rew[indices], dtype=torch.float32)


def target_q_fn_multidim(buffer, indices):
    return target_q_fn(buffer, indices).unsqueeze(1).repeat(1, 51)


@pytest.mark.parametrize("timelimit", [False, True])
def test_nstep_returns(timelimit):
    from tianshou_b200.algorithm import Algorithm
    from tianshou_b200.data import Batch, ReplayBuffer, to_numpy
    buf = ReplayBuffer(10, device=DEV)
    for i in range(12):
        buf.add(Batch(obs=0, act=0, rew=i + 1, terminated=(i % 4 == 3) and not (timelimit and i == 3),
                      truncated=timelimit and i == 3))
    batch, indices = buf.sample(0)
    assert np.allclose(indices, [2, 3, 4, 5, 6, 7, 8, 9, 0, 1])
    for n_step in (1, 2, 10):
        returns = to_numpy(Algorithm.compute_nstep_return(batch, buf, indices, target_q_fn, gamma=0.1, n_step=n_step)
                           .pop("returns").reshape(-1))
        assert np.allclose(returns, NSTEP_KATS[timelimit][n_step])
        multi = to_numpy(Algorithm.compute_nstep_return(batch, buf, indices, target_q_fn_multidim, gamma=0.1,
                                                        n_step=n_step).pop("returns"))
        assert multi.shape == (10, 51)
        assert np.allclose(multi, returns[:, np.newaxis])


def test_nstep_size_mismatch_raises():
    from tianshou_b200.algorithm import Algorithm
    from tianshou_b200.data import Batch, ReplayBuffer
    buf = ReplayBuffer(10, device=DEV)
    for i in range(5):
        buf.add(Batch(obs=0, act=0, rew=1.0, terminated=False, truncated=False))
    batch, indices = buf.sample(0)
    with pytest.raises(ValueError):
        Algorithm.compute_nstep_return(batch, buf, indices[:-1], target_q_fn)


def test_replaybuffermanager_indices():
    """test/base/test_buffer.py:740-972 (index part) on the device kernels."""
    from tianshou_b200.data import Batch, VectorReplayBuffer
    from test_oracle import (MANAGER_KAT_DONE, MANAGER_KAT_NEXT, MANAGER_KAT_NEXT2, MANAGER_KAT_PREV,
                             MANAGER_KAT_PREV2)
    buf = VectorReplayBuffer(20, 4, device=DEV)
    batch = Batch(obs=[1, 2, 3], act=[1, 2, 3], rew=[1, 2, 3], terminated=[0, 0, 1], truncated=[0, 0, 0])
    ptr, ep_rew, ep_len, ep_idx = buf.add(batch, buffer_ids=[0, 1, 2])
    assert np.all(ep_len == [0, 0, 1]) and np.all(ep_rew == [0, 0, 3])
    assert np.all(ptr == [0, 5, 10]) and np.all(ep_idx == [0, 5, 10])
    with pytest.raises(NotImplementedError):
        buf.update(buf)
    indices = buf.sample_indices(11000)
    assert np.bincount(indices)[[0, 5, 10]].min() >= 3000
    batch, indices = buf.sample(0)
    assert np.allclose(indices, [0, 5, 10])
    assert np.allclose(buf.prev(indices), indices) and np.allclose(buf.next(indices), indices)
    assert np.allclose(buf.unfinished_index(), [0, 5])
    buf.add(Batch(obs=[4], act=[4], rew=[4], terminated=[1], truncated=[0]), buffer_ids=[3])
    assert np.allclose(buf.unfinished_index(), [0, 5])
    batch, indices = buf.sample(10)
    batch, indices = buf.sample(0)
    assert np.allclose(indices, [0, 5, 10, 15])
    data = np.array([0, 0, 0, 0])
    buf.add(Batch(obs=data, act=data, rew=data, terminated=data, truncated=data), buffer_ids=[0, 1, 2, 3])
    buf.add(Batch(obs=data, act=data, rew=data, terminated=1 - data, truncated=data), buffer_ids=[0, 1, 2, 3])
    assert len(buf) == 12
    buf.add(Batch(obs=data, act=data, rew=data, terminated=data, truncated=data), buffer_ids=[0, 1, 2, 3])
    buf.add(Batch(obs=data, act=data, rew=data, terminated=[0, 1, 0, 1], truncated=data), buffer_ids=[0, 1, 2, 3])
    assert len(buf) == 20
    indices = buf.sample_indices(120000)
    assert np.bincount(indices).min() >= 5000
    indices = buf.sample_indices(0)
    assert np.allclose(indices, np.arange(len(buf)))
    assert np.allclose(buf.done, MANAGER_KAT_DONE)
    assert np.allclose(buf.prev(indices), MANAGER_KAT_PREV)
    assert np.allclose(buf.next(indices), MANAGER_KAT_NEXT)
    assert np.allclose(buf.unfinished_index(), [4, 14])
    ptr, ep_rew, ep_len, ep_idx = buf.add(Batch(obs=[1], act=[1], rew=[1], terminated=[1], truncated=[0]), buffer_ids=[2])
    assert np.all(ep_len == [3]) and np.all(ep_rew == [1]) and np.all(ptr == [10]) and np.all(ep_idx == [13])
    assert np.allclose(buf.unfinished_index(), [4])
    indices = np.array(sorted(buf.sample_indices(0)))
    assert np.allclose(indices, np.arange(len(buf)))
    assert np.allclose(buf.prev(indices), MANAGER_KAT_PREV2)
    assert np.allclose(buf.next(indices), MANAGER_KAT_NEXT2)
    assert buf.prev(-1) == buf.prev(np.array([buf.maxsize - 1]))[0]
    assert buf.next(-1) == buf.next(np.array([buf.maxsize - 1]))[0]
    assert buf.sample_indices(-1).tolist() == []


def test_single_buffer_indices_vs_reference():
    from tianshou_b200.data import Batch, ReplayBuffer
    from ts_testutil import load_golden
    g = load_golden("index_ref.npz")
    buf = ReplayBuffer(10, device=DEV)
    for i in range(12):
        buf.add(Batch(obs=0, act=0, rew=i + 1, terminated=i % 4 == 3, truncated=False))
    q = np.arange(10)
    assert np.array_equal(buf.prev(q), g["single_prev"]) and np.array_equal(buf.next(q), g["single_next"])
    assert np.array_equal(buf.sample_indices(0), g["single_all"])
    assert np.array_equal(buf.unfinished_index(), g["single_unfinished"])


def test_ignore_obs_next_and_stacking():
    """obs_next through next() and frame stacks through prev() (buffer_base.py:557-649)."""
    from tianshou_b200.data import Batch, ReplayBuffer
    buf = ReplayBuffer(9, stack_num=4, ignore_obs_next=True, device=DEV)
    for i in range(16):
        done = i % 5 == 0
        buf.add(Batch(obs=i, act=i, rew=i, terminated=done, truncated=False))
    idx = buf.sample_indices(0)
    b = buf[idx]
    assert b.obs.shape == (9, 4)
    # last frame of every stack is the slot's own obs; obs_next's last frame is obs[next(idx)]
    assert np.array_equal(b.obs[:, -1], buf.obs[idx])
    assert np.array_equal(b.obs_next[:, -1], buf.obs[buf.next(idx)])


def test_from_data_buffer_indices_vs_reference():
    """``ReplayBuffer.from_data`` leaves ``_insertion_idx`` = 0 and ``last_index`` = 0 with size = N
    (buffer_base.py:382-418): ``sample_indices(0)`` must be ``arange(N)`` -- not the rotation the kernel used to derive
    from ``last_index + 1`` -- and keep following the insertion index through later ``add`` calls.  Golden =
    outputs of the imported reference (``oracle/gen_golden.py fromdata``)."""
    from tianshou_b200.data import Batch, ReplayBuffer
    g = load_golden("fromdata_ref.npz")
    for case in range(int(g["n_cases"])):
        p = f"fd{case}_"
        buf = ReplayBuffer.from_data(g[p + "obs"], g[p + "act"], g[p + "rew"], g[p + "terminated"], g[p + "truncated"],
                                     g[p + "terminated"] | g[p + "truncated"], g[p + "obs_next"])
        for j in range(int(g[p + "extra"])):
            a = p + f"add{j}_"
            buf.add(Batch(obs=g[a + "obs"], act=g[a + "act"], rew=float(g[a + "rew"]), terminated=bool(g[a + "terminated"]),
                          truncated=bool(g[a + "truncated"]), obs_next=g[a + "obs_next"]))
        n = int(g[p + "len"])
        assert len(buf) == n
        assert np.array_equal(buf.sample_indices(0), g[p + "all"]), f"case {case}"
        q = g[p + "query"] % n
        assert np.array_equal(buf.prev(q), g[p + "prev"]) and np.array_equal(buf.next(q), g[p + "next"])
        assert np.array_equal(buf.unfinished_index(), g[p + "unfinished"])
        assert np.array_equal(np.asarray(buf.done, dtype=bool), g[p + "done"])
