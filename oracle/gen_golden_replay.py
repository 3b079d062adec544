"""TEST INFRASTRUCTURE -- replays the buffer construction script of ``gen_golden.gen_index`` with
any Batch / VectorReplayBuffer implementation (no reference import), so host-side bookkeeping can
be compared with the stored reference states on a box without /root/reference."""
from __future__ import annotations

import numpy as np

INDEX_CASES = [(4, 5, 3, 0.3, 4), (4, 5, 12, 0.3, 4), (7, 9, 20, 0.1, 5), (32, 16, 10, 0.05, 7),
               (1, 10, 12, 0.25, 100), (5, 8, 0, 0.1, 5)]


def synth_rollout(rng, E, steps, obs_dim, act_dim, p_term, trunc_len):
    t_in_ep = np.zeros(E, dtype=np.int64)
    obs = rng.standard_normal((E, obs_dim)).astype(np.float32)
    out = []
    for _ in range(steps):
        act = rng.standard_normal((E, act_dim)).astype(np.float32)
        rew = rng.standard_normal(E)
        obs_next = rng.standard_normal((E, obs_dim)).astype(np.float32)
        term = rng.random(E) < p_term
        t_in_ep += 1
        trunc = (t_in_ep >= trunc_len) & ~term
        out.append(dict(obs=obs, act=act, rew=rew, terminated=term, truncated=trunc, obs_next=obs_next))
        done = term | trunc
        t_in_ep[done] = 0
        fresh = rng.standard_normal((E, obs_dim)).astype(np.float32)
        obs = np.where(done[:, None], fresh, obs_next)
    return out


def replay_index_cases(rng, VectorReplayBuffer, Batch):
    for E, cap, steps, p_term, trunc_len in INDEX_CASES:
        buf = VectorReplayBuffer(E * cap, E)
        roll = synth_rollout(rng, E, max(steps, 1), 2, 1, p_term, trunc_len)

        def sub(s, ids):
            return Batch(obs=s["obs"][ids], act=s["act"][ids], rew=s["rew"][ids], terminated=s["terminated"][ids],
                         truncated=s["truncated"][ids], obs_next=s["obs_next"][ids])

        if steps == 0:
            ids = np.array([0, 2])
            buf.add(sub(roll[0], ids), buffer_ids=ids)
        else:
            for s in roll:
                buf.add(sub(s, np.arange(E)), buffer_ids=np.arange(E))
            s = synth_rollout(rng, E, 1, 2, 1, p_term, trunc_len)[0]
            ids = np.arange(0, E, 2)
            buf.add(sub(s, ids), buffer_ids=ids)
        # the generator draws 50 random query indices here
        rng.integers(0, buf.maxsize, 50)
        yield buf, None
