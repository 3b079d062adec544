"""TEST INFRASTRUCTURE -- CPU restatement (plain torch on the host, autograd) of the reference's off-policy update bodies.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; ``tianshou_b200`` never does.
The reference's own implementation of these paths IS torch code; this file restates it without the framework around it
(no Batch / Policy / Collector), each step citing the reference lines (paths relative to /root/reference/tianshou):

  SAC  : algorithm/modelfree/sac.py:108-131 (tanh-Gaussian head), :298-336 (target value, update), modelfree/td3.py:94-102,
         modelfree/ddpg.py:267-285 (critic squared loss), algorithm/algorithm_base.py:721-817 + :1160-1222 (n-step return),
         utils/lagged_network.py:8-18 (Polyak)
  DQN  : algorithm/modelfree/dqn.py:365-404, env/atari/atari_network.py:26-122, data/buffer/buffer_base.py:557-603

PINNING: tests/test_oracle_offpolicy.py replays tests/golden/sac_ref.npz / dqn_ref*.npz (outputs of the imported reference,
oracle/gen_golden_offpolicy.py) through these functions: indices are taken from the golden file, returns / losses / TD errors /
post-update parameters must match.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
from torch import nn

_F32_EPS = float(np.finfo(np.float32).eps)
SIGMA_MIN, SIGMA_MAX = -20.0, 2.0


# ------------------------------------------------------------------------------------------------ n-step return
def nstep_return(rew: np.ndarray, end_flag: np.ndarray, target_q: np.ndarray, indices: np.ndarray, gamma: float, n_step: int) -> np.ndarray:
    """numba ``_nstep_return`` (algorithm_base.py:1160-1222).  ``indices`` = [n_step, I] stacked next-chains."""
    gamma_buffer = np.ones(n_step + 1, dtype=np.float64)
    for i in range(1, n_step + 1):
        gamma_buffer[i] = gamma_buffer[i - 1] * gamma
    target_shape = target_q.shape
    bsz = target_shape[0]
    target_q = target_q.reshape(bsz, -1).astype(np.float64)
    returns = np.zeros(target_q.shape)
    gammas = np.full(indices[0].shape, n_step)
    for n in range(n_step - 1, -1, -1):
        now = indices[n]
        gammas[end_flag[now] > 0] = n + 1
        returns[end_flag[now] > 0] = 0.0
        returns = rew[now].reshape(bsz, 1) + gamma * returns
    target_q = target_q * gamma_buffer[gammas].reshape(bsz, 1) + returns
    return target_q.reshape(target_shape)


def next_index(index: np.ndarray, offset: np.ndarray, done: np.ndarray, last_index: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    """ReplayBufferManager ``_next_index`` (data/buffer/manager.py:339-363)."""
    index = index % offset[-1]
    out = np.zeros_like(index)
    for start, end, cur_len, last in zip(offset[:-1], offset[1:], lengths, last_index):
        mask = (start <= index) & (index < end)
        cur_len = max(1, cur_len)
        if np.sum(mask) > 0:
            sub = index[mask]
            end_flag = done[sub] | (sub == last)
            out[mask] = (sub - start + (1 - end_flag)) % cur_len + start
    return out


def prev_index(index: np.ndarray, offset: np.ndarray, done: np.ndarray, last_index: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    """ReplayBufferManager ``_prev_index`` (data/buffer/manager.py:311-336)."""
    index = index % offset[-1]
    out = np.zeros_like(index)
    for start, end, cur_len, last in zip(offset[:-1], offset[1:], lengths, last_index):
        mask = (start <= index) & (index < end)
        cur_len = max(1, cur_len)
        if np.sum(mask) > 0:
            sub = (index[mask] - start - 1) % cur_len
            end_flag = done[sub + start] | (sub + start == last)
            out[mask] = (sub + end_flag) % cur_len + start
    return out


def compute_nstep_targets(buf: dict, indices: np.ndarray, target_q_fn, gamma: float, n_step: int) -> np.ndarray:
    """``Algorithm.compute_nstep_return`` (algorithm_base.py:772-811).  buf: rew, done, terminated, offset, last_index, lengths."""
    chain = [indices]
    for _ in range(n_step - 1):
        chain.append(next_index(chain[-1], buf["offset"], buf["done"], buf["last_index"], buf["lengths"]))
    chain = np.stack(chain)
    terminal = chain[-1]
    with torch.no_grad():
        tq = target_q_fn(terminal)
    tq = tq.numpy().reshape(len(indices), -1)
    tq = tq * (~buf["terminated"][terminal]).reshape(-1, 1)
    end_flag = buf["done"].copy()
    unfinished = np.array([last for last, n in zip(buf["last_index"], buf["lengths"]) if n > 0 and not buf["done"][last]], dtype=np.int64)
    end_flag[unfinished] = True
    return nstep_return(buf["rew"], end_flag, tq, chain, gamma, n_step).astype(np.float32)


# ------------------------------------------------------------------------------------------------ SAC
def mlp(sizes: list[int], final_act: bool) -> nn.Sequential:
    mods: list[nn.Module] = []
    for i in range(len(sizes) - 1):
        mods.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2 or final_act:
            mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class SacNets:
    """Same parameter order as the reference modules' ``parameters()``: actor = trunk, mu, sigma; critic = trunk, last."""

    def __init__(self, obs: int, act: int, hidden: tuple[int, ...]):
        self.a_trunk = mlp([obs, *hidden], True)
        self.a_mu, self.a_sigma = nn.Linear(hidden[-1], act), nn.Linear(hidden[-1], act)
        self.c = [nn.Sequential(mlp([obs + act, *hidden], True), nn.Linear(hidden[-1], 1)) for _ in range(2)]
        self.c_old = [copy.deepcopy(c) for c in self.c]

    def actor_params(self) -> list[nn.Parameter]:
        return [*self.a_trunk.parameters(), *self.a_mu.parameters(), *self.a_sigma.parameters()]

    def policy(self, obs: torch.Tensor, noise: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """SACPolicy.forward (sac.py:108-131) with the rsample noise given: (squashed action, log_prob [B, 1])."""
        h = self.a_trunk(obs)
        mu = self.a_mu(h)
        sigma = torch.clamp(self.a_sigma(h), min=SIGMA_MIN, max=SIGMA_MAX).exp()       # continuous.py:231-235
        x = mu + sigma * noise                                                          # Normal.rsample
        log_prob = (-((x - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - np.log(np.sqrt(2 * np.pi))).sum(-1, keepdim=True)
        a = torch.tanh(x)
        log_prob = log_prob - torch.log(1 - a.pow(2) + _F32_EPS).sum(-1, keepdim=True)  # sac.py:25-39
        return a, log_prob


def sac_update(nets: SacNets, opts: list[torch.optim.Optimizer], buf: dict, indices: np.ndarray, noise_target: torch.Tensor,
               noise_actor: torch.Tensor, gamma: float, n_step: int, alpha: float, tau: float) -> dict:
    """One ``SAC.update`` on the sampled ``indices`` (sac.py:298-336 + ddpg.py:267-302)."""
    def target_q(terminal: np.ndarray) -> torch.Tensor:
        obs_next = torch.from_numpy(buf["obs_next"][terminal])
        a, lp = nets.policy(obs_next, noise_target)
        x = torch.cat([obs_next, a], dim=1)
        return torch.min(nets.c_old[0](x), nets.c_old[1](x)) - alpha * lp

    returns = torch.from_numpy(compute_nstep_targets(buf, indices, target_q, gamma, n_step)).flatten()
    obs, act = torch.from_numpy(buf["obs"][indices]), torch.from_numpy(buf["act"][indices])
    losses, tds = [], []
    for k in range(2):
        q = nets.c[k](torch.cat([obs, act], dim=1)).flatten()
        td = q - returns
        loss = td.pow(2).mean()
        opts[1 + k].zero_grad(); loss.backward(); opts[1 + k].step()
        losses.append(float(loss)); tds.append(td.detach())
    a, lp = nets.policy(obs, noise_actor)
    x = torch.cat([obs, a], dim=1)
    actor_loss = (alpha * lp.flatten() - torch.min(nets.c[0](x).flatten(), nets.c[1](x).flatten())).mean()
    opts[0].zero_grad(); actor_loss.backward(); opts[0].step()
    with torch.no_grad():
        for k in range(2):
            for t, s in zip(nets.c_old[k].parameters(), nets.c[k].parameters(), strict=True):
                t.copy_(tau * s + (1 - tau) * t)
    return dict(returns=returns.numpy(), actor_loss=float(actor_loss), critic1_loss=losses[0], critic2_loss=losses[1],
                weight=((tds[0] + tds[1]) / 2.0).numpy())


# ------------------------------------------------------------------------------------------------ DQN
def nature_cnn(c: int, h: int, w: int, actions: int) -> nn.Sequential:
    """DQNet (atari_network.py:77-96)."""
    conv = nn.Sequential(nn.Conv2d(c, 32, 8, 4), nn.ReLU(), nn.Conv2d(32, 64, 4, 2), nn.ReLU(), nn.Conv2d(64, 64, 3, 1), nn.ReLU(),
                         nn.Flatten())
    with torch.no_grad():
        feat = int(np.prod(conv(torch.zeros(1, c, h, w)).shape[1:]))
    return nn.Sequential(conv, nn.Linear(feat, 512), nn.ReLU(), nn.Linear(512, actions))


def stacked_obs(frames: np.ndarray, index: np.ndarray, buf: dict, stack: int) -> np.ndarray:
    """ReplayBuffer.get frame stacking (buffer_base.py:585-600): [..., prev(prev(i)), prev(i), i] along axis 1."""
    out, cur = [], index
    for _ in range(stack):
        out.insert(0, frames[cur])
        cur = prev_index(cur, buf["offset"], buf["done"], buf["last_index"], buf["lengths"])
    return np.stack(out, axis=1)


def dqn_update(net: nn.Module, net_old: nn.Module | None, opt: torch.optim.Optimizer, buf: dict, indices: np.ndarray,
               is_weight: np.ndarray | None, gamma: float, n_step: int, is_double: bool, huber: float | None, denom: float = 255.0,
               stack: int = 4, sync_target: bool = False) -> dict:
    """One ``DQN.update`` on ``indices`` (dqn.py:365-404).  buf["obs"] = uint8 frames.  ``sync_target``: this is an iteration
    on which the lagged network is refreshed -- the reference does that at the top of ``_update_with_batch`` (:386), i.e. AFTER
    ``_preprocess_batch`` computed the n-step targets with the old copy (algorithm_base.py:619-623)."""
    def q_of(model: nn.Module, idx: np.ndarray) -> torch.Tensor:
        x = stacked_obs(buf["obs"], idx, buf, stack) / denom                 # numpy f64 (atari_network.py:48-55)
        return model(torch.as_tensor(x, dtype=torch.float32))

    def target_q(terminal: np.ndarray) -> torch.Tensor:
        nxt = next_index(terminal, buf["offset"], buf["done"], buf["last_index"], buf["lengths"])      # obs_next = obs[next(i)]
        q_online = q_of(net, nxt)
        q_tgt = q_of(net_old, nxt) if net_old is not None else q_online
        if is_double:
            return q_tgt[np.arange(len(nxt)), q_online.argmax(dim=1)]
        return q_tgt.max(dim=1)[0]

    returns = torch.from_numpy(compute_nstep_targets(buf, indices, target_q, gamma, n_step)).flatten()
    if sync_target and net_old is not None:
        net_old.load_state_dict(net.state_dict())
    q = q_of(net, indices)
    q = q[np.arange(len(indices)), buf["act"][indices]]
    td = returns - q
    if huber is not None:
        loss = torch.nn.functional.huber_loss(q.reshape(-1, 1), returns.reshape(-1, 1), delta=huber, reduction="mean")
    else:
        w = torch.as_tensor(is_weight, dtype=torch.float32) if is_weight is not None else 1.0
        loss = (td.pow(2) * w).mean()
    opt.zero_grad(); loss.backward(); opt.step()
    return dict(returns=returns.numpy(), td=td.detach().numpy(), loss=float(loss))
